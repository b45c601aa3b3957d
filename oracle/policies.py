"""TEST INFRASTRUCTURE (oracle): NumPy restatement of the reference's rule-based benchmark policies, per env.

Follows fleetrl/benchmarking/uncontrolled_charging.py:51-54, distributed_charging.py:50-54,
night_charging.py:81-98 and FleetEnv.get_dist_factor (fleet_environment.py:782-799).  Only tests/ may import this.
The reference scripts pass float64 actions to VecEnv.step; the product emits float32 (the action space's dtype), so
the comparison casts to float32.
"""
import numpy as np


def dist_factor(consts, tables, t, target_soc):
    """hours_needed / (hours_left + 0.001) of the SCHEDULE observation at table index t, for one env: [N] float64."""
    there = np.asarray(tables["there"])[:, t].astype(np.float64)
    cl = target_soc * there - np.asarray(tables["soc_on_return"])[:, t]                     # observer_bl_pv.py:86-88
    hn = cl * consts.lc_batt_cap / (consts.evse_max_power * consts.charging_eff)            # :89
    return np.divide(hn, np.add(np.asarray(tables["time_left"])[:, t], 0.001))


def uncontrolled(n_evs):
    return np.ones(n_evs)


def distributed(consts, tables, t, target_soc):
    return np.clip(np.multiply(np.ones(consts.num_evs), dist_factor(consts, tables, t, target_soc)), 0, 1)


class NightPolicy:
    """The loop body of night_charging.py:81-98 for one env; `charging` / `charging_start` persist like its locals."""

    def __init__(self, consts, tables, charging_hour, charging_minute, max_time_needed_int):
        self.c, self.tb = consts, tables
        self.ch, self.cm, self.max_h = charging_hour, charging_minute, max_time_needed_int
        self.charging = False
        self.charging_start = 0

    def actions(self, t, target_soc):
        c = self.c
        hour, minute = int(self.tb["hour"][t]), int(self.tb["minute"][t])
        if c.is_caretaker and 11 <= hour <= 14:
            return distributed(c, self.tb, t, target_soc)
        if ((self.ch <= hour) and (self.cm <= minute)) or self.charging:
            if not self.charging:
                self.charging_start = t
            self.charging = True
            a = np.ones(c.num_evs)
        else:
            a = np.zeros(c.num_evs)
        if self.charging and ((t - self.charging_start) * c.dt > self.max_h):          # (time - charging_start) in hours
            self.charging = False
        return a
