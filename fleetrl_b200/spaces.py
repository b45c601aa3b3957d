"""Minimal `Box` space (gymnasium is optional): stores low/high/shape/dtype like gymnasium.spaces.Box and is what
FleetEnv.observation_space / action_space return (fleet_environment.py:316-325).  If gymnasium is importable its
Box is used instead so that SB3's space checks accept it."""
import numpy as np

try:  # pragma: no cover - depends on the installation
    from gymnasium.spaces import Box as _GymBox
except Exception:  # gymnasium is not installed in the build image
    _GymBox = None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def Box(low, high, shape=None, dtype=np.float32):
    if _GymBox is not None:
        return _GymBox(low=low, high=high, shape=shape, dtype=dtype)
    return _Box(low, high, shape, dtype)


def observation_box(dim, normalized):
    """make_boundaries: OracleNormalization -> [0,1] float32 (oracle_normalization.py:164-173),
    UnitNormalization -> (-inf, inf) (unit_normalization.py:23-31)."""
    if normalized:
        return Box(np.zeros(dim, np.float32), np.ones(dim, np.float32), dtype=np.float32)
    return Box(np.full(dim, -np.inf, np.float32), np.full(dim, np.inf, np.float32), dtype=np.float32)


def action_box(num_cars):
    """fleet_environment.py:322-325"""
    return Box(low=-1, high=1, shape=(num_cars,), dtype=np.float32)
