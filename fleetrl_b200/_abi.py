"""ctypes mirror of include/fleetstep.h (FleetConsts, FleetTables, enums).

Pure declarations — no computation lives here.  Used by the product binding (fleetrl_b200/_lib.py) and, for
the struct layouts only, by the test-side oracle wrapper (oracle/oracle.py).
"""
import ctypes as C

import numpy as np

ABI_VERSION = 2

FLEET_DEG_SEI = 0
FLEET_DEG_EMPIRICAL = 1


class FleetConsts(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("num_evs", C.c_int32), ("table_len", C.c_int32),
        ("steps_per_hour", C.c_int32), ("episode_steps", C.c_int32), ("price_lookahead", C.c_int32),
        ("bl_pv_lookahead", C.c_int32), ("include_price", C.c_int32), ("include_building", C.c_int32),
        ("include_pv", C.c_int32), ("aux", C.c_int32), ("normalize", C.c_int32), ("is_caretaker", C.c_int32),
        ("calc_degradation", C.c_int32), ("deg_mode", C.c_int32), ("carry_degradation_state", C.c_int32),
        ("auto_reset", C.c_int32), ("start_lo", C.c_int32), ("start_hi", C.c_int32),
        ("rf_ring_rows", C.c_int32), ("rf_stack_depth", C.c_int32), ("reserved0", C.c_int32),
        ("seed", C.c_uint64),
        ("dt", C.c_double),
        ("init_battery_cap", C.c_double), ("obc_max_power", C.c_double), ("charging_eff", C.c_double),
        ("discharging_eff", C.c_double), ("def_soc", C.c_double), ("temperature", C.c_double),
        ("target_soc", C.c_double), ("target_soc_lunch", C.c_double), ("min_laxity", C.c_double),
        ("fixed_markup", C.c_double), ("variable_multiplier", C.c_double), ("feed_in_deduction", C.c_double),
        ("evse_max_power", C.c_double), ("grid_connection", C.c_double), ("lc_batt_cap", C.c_double),
        ("price_multiplier", C.c_double), ("fully_charged_reward", C.c_double),
        ("penalty_invalid_action", C.c_double), ("penalty_overcharging", C.c_double),
        ("penalty_overloading", C.c_double), ("clip_overcharging", C.c_double),
        ("init_soh", C.c_double), ("soc_eps", C.c_double),
        ("max_time_left", C.c_double), ("min_price", C.c_double), ("max_price", C.c_double),
        ("min_tariff", C.c_double), ("max_tariff", C.c_double), ("max_building", C.c_double),
        ("max_pv", C.c_double),
    ]

    INT_FIELDS = ("num_evs table_len steps_per_hour episode_steps price_lookahead bl_pv_lookahead include_price "
                  "include_building include_pv aux normalize is_caretaker calc_degradation deg_mode "
                  "carry_degradation_state auto_reset start_lo start_hi rf_ring_rows rf_stack_depth seed").split()

    def to_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}

    @classmethod
    def from_dict(cls, d):
        c = cls()
        for name, _ in cls._fields_:
            if name in d:
                v = d[name]
                setattr(c, name, int(v) if name in cls.INT_FIELDS or name in ("abi_version", "reserved0") else float(v))
        c.abi_version = ABI_VERSION
        return c


class FleetTables(C.Structure):
    _fields_ = [
        ("there", C.c_void_p), ("time_left", C.c_void_p), ("soc_on_return", C.c_void_p),
        ("delu", C.c_void_p), ("tariff", C.c_void_p), ("load", C.c_void_p), ("pv", C.c_void_p),
        ("price_reward_curve", C.c_void_p), ("tariff_reward_curve", C.c_void_p),
        ("cal_sincos", C.c_void_p), ("hour", C.c_void_p), ("minute", C.c_void_p),
    ]

    DTYPES = {"there": np.uint8, "time_left": np.float64, "soc_on_return": np.float64, "delu": np.float64,
              "tariff": np.float64, "load": np.float64, "pv": np.float64, "price_reward_curve": np.float64,
              "tariff_reward_curve": np.float64, "cal_sincos": np.float64, "hour": np.uint8, "minute": np.uint8}

    @classmethod
    def from_arrays(cls, arrays: dict):
        """arrays: name -> ndarray or None.  Returns (struct, keepalive list of contiguous arrays)."""
        t = cls()
        keep = []
        for name, _ in cls._fields_:
            a = arrays.get(name)
            if a is None:
                setattr(t, name, None)
                continue
            a = np.ascontiguousarray(a, dtype=cls.DTYPES[name])
            keep.append(a)
            setattr(t, name, a.ctypes.data)
        return t, keep


# FLEET_F_* : name -> (id, numpy dtype, per_ev)
FIELDS = {
    "soc": (0, np.float64, True), "hours_left": (1, np.float32, True), "soc_deg": (2, np.float64, True),
    "soh": (3, np.float64, True), "target_soc": (4, np.float64, True), "time_idx": (5, np.int32, False),
    "finish_idx": (6, np.int32, False), "reward64": (7, np.float64, False), "cashflow": (8, np.float64, False),
    "rf_len": (9, np.int32, True), "fd_cyc": (10, np.float64, True), "life": (11, np.float64, True),
    "ep_return": (12, np.float64, False), "ep_count": (13, np.int32, False),
    "last_ep_return": (14, np.float64, False), "n_cycles": (15, np.int32, True),
    "last_deg": (16, np.float64, True), "overload": (17, np.float64, False), "soc_viol": (18, np.float64, False),
    "charge_log": (19, np.float64, True),        # only after FleetStepHandle.enable_charge_log()
}

# FLEET_S_*
STATS = ["episodes", "ep_return", "steps", "reward", "cashflow", "penalty", "overload_kw", "soc_viol", "n_viol",
         "degradation"]
