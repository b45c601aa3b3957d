"""fleetrl_b200 — the FleetRL environment step (FleetEnv.reset/step) as hand-written sm_100a CUDA behind a C ABI.

Only what the path needs lives here: csrc/ (kernels + C ABI), the ctypes binding, and the host-side mirror of
the reference's FleetEnv / SB3 VecEnv interface.
"""
__version__ = "0.1.0"
