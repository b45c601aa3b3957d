"""Empty stand-in for matplotlib: the reference imports pyplot/patches at module load
(fleet_environment.py:8-9, rendering/render.py:1-2) but the env step never draws."""
