"""Make the UNMODIFIED reference package (/root/reference/fleetrl) importable in the build container.

TEST INFRASTRUCTURE ONLY — used by oracle/gen_golden.py and by the container-only tests that
compare against the live reference.  Nothing in fleetrl_b200/ imports this.  The GPU box has no
/root/reference, so nothing marked `gpu`, `smoke()` or `bench.py` may depend on it.

What it does:
  1. puts stub `gymnasium`, `matplotlib`, `rainflow` packages (this directory) and /root/reference on sys.path;
  2. pandas-3 compatibility: the reference is written for pandas 2.2 and uses the removed offset aliases
     "15T", "1H", "H", "M" (data_processing.py:58,75-77,392,406; observer_*.py resample("H");
     random_time_picker.py:25) and integer-positional `Series[0]` on a labelled index
     (data_processing.py:422,428-429);
  3. pandas-3 Copy-on-Write turns `df.loc[:, "time_left"].fillna(0, inplace=True)`
     (data_processing.py:210) into a no-op; under pandas 2.2 it fills.  `install()` wraps
     `DataLoader.compute_from_schedule` to apply the pandas-2.2 result (NaN -> 0) afterwards, so the
     reference behaves as on its pinned dependency set.
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("FLEETRL_REFERENCE_ROOT", "/root/reference")
_ALIASES = {"15T": "15min", "1H": "1h", "H": "h", "M": "ME", "T": "min", "30T": "30min", "5T": "5min"}
_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fleetrl"))


def _fix(freq):
    return _ALIASES.get(freq, freq) if isinstance(freq, str) else freq


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (REFERENCE_ROOT, here):
        if p not in sys.path:
            sys.path.insert(0, p)

    import pandas as pd
    from pandas.core.groupby import DataFrameGroupBy

    warnings.filterwarnings("ignore")

    def wrap_resample(cls):
        orig = cls.resample

        def resample(self, rule, *a, **k):
            return orig(self, _fix(rule), *a, **k)

        cls.resample = resample

    for cls in (pd.DataFrame, pd.Series, DataFrameGroupBy):
        wrap_resample(cls)

    orig_date_range = pd.date_range

    def date_range(*a, **k):
        if "freq" in k:
            k["freq"] = _fix(k["freq"])
        return orig_date_range(*a, **k)

    pd.date_range = date_range

    orig_getitem = pd.Series.__getitem__

    def getitem(self, key):
        try:
            return orig_getitem(self, key)
        except KeyError:
            if isinstance(key, int):
                return self.iloc[key]
            raise

    pd.Series.__getitem__ = getitem

    from fleetrl.utils.data_processing.data_processing import DataLoader

    orig_compute = DataLoader.compute_from_schedule

    def compute_from_schedule(self, ev_conf, time_conf, target_soc):
        orig_compute(self, ev_conf, time_conf, target_soc)
        # pandas-2.2 semantics of data_processing.py:210 (see module docstring, item 3)
        self.schedule["time_left"] = self.schedule["time_left"].fillna(0)

    DataLoader.compute_from_schedule = compute_from_schedule
    _installed = True


def base_config(**over):
    """A complete env_config dict (every key fleet_environment.py:129-211,243,285 reads with [])."""
    cfg = {
        "data_path": os.path.join(REFERENCE_ROOT, "inputs"),
        "use_case": "lmd",
        "schedule_name": "1_lmd.csv",
        "building_name": "load_lmd.csv",
        "pv_name": None,
        "price_name": "spot_2020_new.csv",
        "tariff_name": "spot_2020_new_tariff.csv",
        "seed": 42,
        "include_building": True, "include_pv": True, "include_price": True,
        "time_picker": "static",
        "max_batt_cap_in_all_use_cases": 60,
        "init_soh": 1.0,
        "log_data": False, "deg_emp": False, "calculate_degradation": True,
        "verbose": 0, "normalize_in_env": False, "aux": True,
        "ignore_price_reward": False, "ignore_overloading_penalty": False,
        "ignore_invalid_penalty": False, "ignore_overcharging_penalty": False,
        "gen_schedule": False, "gen_start_date": "2020-01-01 00:00", "gen_end_date": "2020-12-30 23:59",
        "gen_name": "gen.csv", "gen_n_evs": 1,
        "spot_markup": None, "spot_mul": None, "feed_in_ded": None,
        "real_time": False,
        "episode_length": 24, "target_soc": 0.85,
    }
    cfg.update(over)
    return cfg


def make_reference_env(cfg, start_time=None, linear_degradation_patch=False):
    """Construct the reference FleetEnv.  `start_time` replaces the time picker;
    `linear_degradation_patch` applies the one-line wiring fix for deg_emp=True (SURVEY B-1)."""
    install()
    from fleetrl.fleet_env.fleet_environment import FleetEnv
    from fleetrl.utils.time_picker.static_time_picker import StaticTimePicker

    env = FleetEnv(cfg)
    if start_time is not None:
        env.time_picker = StaticTimePicker(start_time=start_time)
    if linear_degradation_patch:
        env.sei_deg = env.emp_deg
    return env
