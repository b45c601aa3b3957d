"""Host side of the rule-based baseline policies (the callers on the other side of the step path).

Mirrors the reference's benchmarking scripts; the actions themselves are computed on the device by
`fleet_policy_actions` (include/fleetstep.h):
  Uncontrolled          fleetrl/benchmarking/uncontrolled_charging.py:51-54
  DistributedCharging   fleetrl/benchmarking/distributed_charging.py:50-54 (FleetEnv.get_dist_factor, fleet_environment.py:782-799)
  NightCharging         fleetrl/benchmarking/night_charging.py:53-98
This module only derives the night policy's window parameters from the schedule, like night_charging.py:53-71.
"""
import math
from dataclasses import dataclass

import numpy as np


@dataclass
class NightParams:
    charging_hour: int
    charging_minute: int
    max_hours: int              # int(max_time_needed): the window closes once it has been open for MORE than this
    max_time_needed: float
    earliest_dep: float


def night_params(built) -> NightParams:
    """night_charging.py:53-71.  The reference finds departures as Location 'home' -> 'driving' rows; the built tables
    carry `there` (PowerRating_kW != 0), whose 1 -> 0 edges are the same rows for schedules in the reference schema."""
    c, tb = built.consts, built.tables
    there = np.asarray(tb["there"])
    dep = (there[:, :-1] == 1) & (there[:, 1:] == 0)             # leaves at column t+1
    cols = np.nonzero(dep.any(axis=0))[0] + 1
    if len(cols) == 0:
        raise ValueError("the schedule has no departures")
    hour = np.asarray(tb["hour"], dtype=np.int64)[cols]
    minute = np.asarray(tb["minute"], dtype=np.int64)[cols]
    tod = hour * 60 + minute
    k = int(np.argmin(tod))
    earliest_dep = hour[k] + minute[k] / 60
    max_time_needed = c.target_soc * c.init_battery_cap / c.charging_eff / c.evse_max_power      # :61
    starting_time = 24 + (earliest_dep - max_time_needed)
    if starting_time > 24:
        starting_time = 23.99
    frac, whole = math.modf(starting_time)
    minutes = np.asarray([0, 15, 30, 45])
    charging_minute = int(minutes[np.abs(minutes - int(frac * 60)).argmin()])
    return NightParams(int(whole), charging_minute, int(max_time_needed), float(max_time_needed), float(earliest_dep))
