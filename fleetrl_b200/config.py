"""Host-side mirror of the reference's configuration objects for the env-step path.

Mirrors (same names, defaults and override order):
  EvConfig     fleetrl/fleet_env/config/ev_config.py:6-18
  ScoreConfig  fleetrl/fleet_env/config/score_config.py:11-24
  TimeConfig   fleetrl/fleet_env/config/time_config.py:10-24
  LoadCalculation._import_company   fleetrl/utils/load_calculation/load_calculation.py:15-60
  FleetEnv.__init__ override order  fleetrl/fleet_env/fleet_environment.py:129-211, 1041-1078
Nothing here touches the GPU; `resolve()` turns an env_config dict into the scalar half of a FleetConsts.
"""
import json
import os
from dataclasses import dataclass, field

MANDATORY_KEYS = (  # accessed with [] in FleetEnv.__init__ (fleet_environment.py:129-211,243,285)
    "seed", "include_price", "include_building", "include_pv", "aux", "normalize_in_env", "data_path", "gen_schedule",
    "schedule_name", "gen_name", "gen_start_date", "gen_end_date", "gen_n_evs", "price_name", "tariff_name",
    "building_name", "pv_name", "use_case", "spot_markup", "spot_mul", "feed_in_ded", "max_batt_cap_in_all_use_cases",
    "episode_length", "target_soc", "ignore_price_reward", "ignore_overloading_penalty", "ignore_invalid_penalty",
    "ignore_overcharging_penalty", "verbose", "calculate_degradation", "log_data", "time_picker", "init_soh",
    "real_time", "deg_emp",
)

_FREQ_MINUTES = {"15T": 15, "15min": 15, "1H": 60, "1h": 60, "H": 60, "h": 60, "30T": 30, "30min": 30, "5T": 5, "5min": 5,
                 "T": 1, "min": 1}


@dataclass
class EvConfig:
    init_battery_cap: float = 60.0
    obc_max_power: float = 100.0
    charging_eff: float = 0.91
    discharging_eff: float = 0.91
    def_soc: float = 0.5
    temperature: float = 25.0
    target_soc: float = 0.85
    target_soc_lunch: float = 0.65
    min_laxity: float = 2
    fixed_markup: float = 10
    variable_multiplier: float = 1.5
    feed_in_deduction: float = 0.25

    @classmethod
    def from_config(cls, cfg):
        return cls(**{k: cfg.get(k, getattr(cls, k)) for k in cls.__dataclass_fields__})


@dataclass
class ScoreConfig:
    price_multiplier: float = 3.33
    price_exponent: float = 1
    fully_charged_reward: float = 1
    penalty_invalid_action: float = -0.2
    penalty_overcharging: float = -0.0055
    penalty_overloading: float = 1
    clip_overcharging: float = -0.2

    @classmethod
    def from_config(cls, cfg):
        return cls(**{k: cfg.get(k, getattr(cls, k)) for k in cls.__dataclass_fields__})


@dataclass
class TimeConfig:
    episode_length: int = 24
    end_cutoff: int = 60
    price_lookahead: int = 8
    bl_pv_lookahead: int = 4
    freq: str = "15T"
    minutes: int = 15
    time_steps_per_hour: int = 4
    dt: float = field(init=False, default=0.25)

    def __post_init__(self):
        self.dt = self.minutes / 60                                   # time_config.py:24

    @classmethod
    def from_config(cls, cfg):
        keys = ("episode_length", "end_cutoff", "price_lookahead", "bl_pv_lookahead", "freq", "minutes", "time_steps_per_hour")
        return cls(**{k: cfg.get(k, getattr(cls, k)) for k in keys})


@dataclass
class Company:
    """LoadCalculation constants per use case (load_calculation.py:32-52)."""
    use_case: str
    evse_max_power: float
    batt_cap: float
    grid_connection: float


def read_config(env_config):
    """fleet_environment.py:121-126 — accept a dict or a path to a JSON file."""
    assert (env_config.__class__ == dict) or (env_config.__class__ == str), 'Invalid config type.'
    if env_config.__class__ == str:
        assert os.path.isfile(env_config), f'Config file not found at {env_config}.'
        with open(env_config, "r") as f:
            return json.load(f)
    return env_config


def init_battery_cap_for(use_case, cfg):
    """specify_company_and_battery_size, fleet_environment.py:1041-1060."""
    if use_case == "ct":
        return 16.7
    if use_case == "ut":
        return 50.0
    if use_case == "lmd":
        return 60.0
    if use_case == "custom":
        return cfg["custom_ev_battery_size_in_kwh"]
    raise TypeError("Company not recognised.")


def company_for(use_case, cfg, max_load, num_cars) -> Company:
    """LoadCalculation._import_company, load_calculation.py:15-60."""
    if use_case == "lmd":
        evse = 11
        grid = max(max_load * 1.1, max_load + 0.5 * num_cars * evse)
        cap = 60
    elif use_case == "ut":
        evse = 22
        grid = max(max_load * 1.1, max_load + 0.5 * num_cars * evse)
        if num_cars > 1:
            grid = 1000
        cap = 50
    elif use_case == "ct":
        evse = 4.6
        grid = max(max_load * 1.1, max_load + 0.5 * num_cars * evse)
        cap = 16.7
    elif use_case == "custom":
        evse = cfg.get("custom_ev_charger_power_in_kw", 120)
        grid = cfg.get("custom_grid_connection_in_kw", 500)
        cap = cfg.get("custom_ev_battery_size_in_kwh", 60)
    else:
        raise TypeError("Company not recognised.")
    return Company(use_case, float(evse), float(cap), float(grid))


@dataclass
class ResolvedConfig:
    cfg: dict
    ev: EvConfig
    score: ScoreConfig
    time: TimeConfig
    use_case: str


def resolve(env_config) -> ResolvedConfig:
    """Apply FleetEnv.__init__'s override sequence (fleet_environment.py:129-211) to the three config objects."""
    cfg = read_config(env_config)
    missing = [k for k in MANDATORY_KEYS if k not in cfg]
    if missing:
        raise KeyError(f"env_config is missing mandatory keys (the reference reads them with []): {missing}")
    time = TimeConfig.from_config(cfg)
    ev = EvConfig.from_config(cfg)
    score = ScoreConfig.from_config(cfg)
    use_case = cfg["use_case"]
    ev.init_battery_cap = init_battery_cap_for(use_case, cfg)                       # :178
    if cfg["spot_markup"] is not None:                                              # change_markups :1062-1068
        ev.fixed_markup = cfg["spot_markup"]
    if cfg["spot_mul"] is not None:
        ev.variable_multiplier = cfg["spot_mul"]
    if cfg["feed_in_ded"] is not None:
        ev.feed_in_deduction = cfg["feed_in_ded"]
    score.price_multiplier = score.price_multiplier * (cfg["max_batt_cap_in_all_use_cases"] / ev.init_battery_cap)  # :194
    time.episode_length = cfg["episode_length"]                                     # :198
    ev.target_soc = cfg["target_soc"]                                               # :199
    if cfg["ignore_price_reward"]:                                                  # adjust_score_config :1070-1078
        score.price_multiplier = 0
    if cfg["ignore_overloading_penalty"]:
        score.penalty_overloading = 0
    if cfg["ignore_invalid_penalty"]:
        score.penalty_invalid_action = 0
    if cfg["ignore_overcharging_penalty"]:
        score.penalty_overcharging = 0
    if cfg["real_time"]:
        raise NotImplementedError("real_time=True (EventManager loop, variable dt) is out of scope; the reference "
                                  "documents it as experimental (fleet_environment.py:692-699)")
    if not cfg["include_price"]:
        raise NotImplementedError("include_price=False is unsupported: the reference raises "
                                  "KeyError('price_reward_curve') in EvCharger.charge (ev_charger.py:155)")
    if time.freq not in _FREQ_MINUTES or _FREQ_MINUTES[time.freq] != time.minutes:
        raise ValueError(f"freq={time.freq!r} and minutes={time.minutes} disagree or are unsupported")
    import struct
    dt = time.minutes / 60
    if struct.unpack("f", struct.pack("f", dt))[0] != dt:
        raise ValueError(f"minutes={time.minutes}: the step length {dt} h is not exactly representable in float32 (the device "
                         "keeps hours_left in float32, exact for 15 / 30 / 60-minute grids); use one of those resolutions")
    return ResolvedConfig(cfg=cfg, ev=ev, score=score, time=time, use_case=use_case)


def default_config(use_case="lmd", **over):
    """A complete env_config dict with the reference's documented keys (fleet_environment.py:81-115) and the values
    of its shipped config.json, for callers that bring their inputs in memory (synthetic fleets, benchmarks)."""
    cfg = {
        "data_path": None, "use_case": use_case, "schedule_name": None, "building_name": None, "pv_name": None,
        "price_name": None, "tariff_name": None, "seed": 42,
        "include_building": True, "include_pv": True, "include_price": True, "time_picker": "random",
        "max_batt_cap_in_all_use_cases": 60, "init_soh": 1.0, "log_data": False, "deg_emp": False,
        "calculate_degradation": True, "verbose": 0, "normalize_in_env": False, "aux": True,
        "ignore_price_reward": False, "ignore_overloading_penalty": False, "ignore_invalid_penalty": False,
        "ignore_overcharging_penalty": False, "gen_schedule": False, "gen_start_date": "2020-01-01 00:00",
        "gen_end_date": "2020-12-30 23:59", "gen_name": "gen.csv", "gen_n_evs": 1, "spot_markup": None,
        "spot_mul": None, "feed_in_ded": None, "real_time": False, "episode_length": 24, "target_soc": 0.85,
    }
    cfg.update(over)
    return cfg
