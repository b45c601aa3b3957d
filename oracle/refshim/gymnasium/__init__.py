"""Minimal stand-in for `gymnasium` (0.29.1 pinned by the reference, not installed here).

TEST INFRASTRUCTURE ONLY. It exists so the unmodified reference package under /root/reference
can be imported in the build container to generate golden vectors (oracle/gen_golden.py).
Only the two names the reference touches are provided: `Env` (fleet_environment.py:50) and
`spaces.Box` (fleet_environment.py:316-325).
"""
from . import spaces  # noqa: F401


class Env:
    def __init__(self, *args, **kwargs):
        pass
