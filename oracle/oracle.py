"""ctypes wrapper around oracle/libfleet_oracle.so (the C restatement in fleet_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from fleetrl_b200._abi import FIELDS, STATS, FleetConsts, FleetTables

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libfleet_oracle.so")
    src = os.path.join(_HERE, "fleet_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "fleetstep.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(FleetConsts), C.POINTER(FleetTables), C.c_int32, C.c_int64]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_obs_dim.argtypes = [C.c_void_p]
        L.oracle_obs_dim.restype = C.c_int32
        L.oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_step_mt.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int32]
        L.oracle_set_next_start.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_set_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_reset_stats.argtypes = [C.c_void_p]
        L.oracle_err_flags.argtypes = [C.c_void_p]
        L.oracle_err_flags.restype = C.c_uint32
        L.oracle_rainflow.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_rainflow.restype = C.c_int32
        _LIB = L
    return _LIB


def rainflow_cycles(series):
    x = np.ascontiguousarray(series, dtype=np.float64)
    out = np.zeros((max(len(x), 1) + 2, 5))
    m = lib().oracle_rainflow(x.ctypes.data, len(x), out.ctypes.data)
    return [(r[0], r[1], r[2], int(r[3]), int(r[4])) for r in out[:m]]


def _ptr(a):
    return None if a is None else a.ctypes.data


class OracleFleet:
    """E independent reference-semantics envs on the CPU.  consts: FleetConsts; tables: dict name->ndarray."""

    def __init__(self, consts: FleetConsts, tables: dict, num_envs: int, env_id_offset: int = 0, threads: int = 1):
        self.consts = consts
        self._tables, self._keep = FleetTables.from_arrays(tables)
        self.E, self.N = int(num_envs), int(consts.num_evs)
        self.threads = threads
        self._h = lib().oracle_create(C.byref(consts), C.byref(self._tables), self.E, env_id_offset)
        if not self._h:
            raise ValueError("oracle_create rejected the configuration")
        self.D = lib().oracle_obs_dim(self._h)
        self._next = None

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, start_idx=None, mask=None):
        obs = np.zeros((self.E, self.D), np.float32)
        s = None if start_idx is None else np.ascontiguousarray(start_idx, np.int32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().oracle_reset(self._h, _ptr(s), _ptr(m), obs.ctypes.data)
        return obs

    def set_next_start(self, next_start):
        self._next = None if next_start is None else np.ascontiguousarray(next_start, np.int32)
        lib().oracle_set_next_start(self._h, _ptr(self._next))

    def step(self, actions, want_terminal=False):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.E, self.N)
        obs = np.zeros((self.E, self.D), np.float32)
        rew = np.zeros(self.E, np.float64)
        cash = np.zeros(self.E, np.float64)
        done = np.zeros(self.E, np.uint8)
        term = np.zeros((self.E, self.D), np.float32) if want_terminal else None
        lib().oracle_step_mt(self._h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, cash.ctypes.data,
                             done.ctypes.data, _ptr(term), self.threads)
        if want_terminal:
            return obs, rew, cash, done, term
        return obs, rew, cash, done

    def step_noout(self, actions):
        """Step without materialising outputs (CPU-baseline timing)."""
        a = np.ascontiguousarray(actions, np.float32)
        lib().oracle_step_mt(self._h, a.ctypes.data, None, None, None, None, None, self.threads)

    def get(self, name):
        fid, dt, per_ev = FIELDS[name]
        out = np.zeros((self.E, self.N) if per_ev else (self.E,), dt)
        if lib().oracle_get_state(self._h, fid, out.ctypes.data) != 0:
            raise KeyError(name)
        return out

    def set(self, name, value):
        fid, dt, per_ev = FIELDS[name]
        v = np.ascontiguousarray(value, dt)
        if lib().oracle_set_state(self._h, fid, v.ctypes.data) != 0:
            raise KeyError(name)

    def stats(self):
        out = np.zeros(len(STATS))
        lib().oracle_get_stats(self._h, out.ctypes.data)
        return dict(zip(STATS, out))

    def err_flags(self):
        return int(lib().oracle_err_flags(self._h))
