"""Baseline-policy host logic vs the live reference (build container only): FleetEnv.get_dist_factor along a trajectory
and the night-charging window parameters of benchmarking/night_charging.py:53-71 evaluated on the reference's own db."""
import math
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.reference

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "refshim"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))


@pytest.mark.parametrize("over", [dict(), dict(use_case="ct", schedule_name="1_ct.csv", building_name="load_ct.csv")])
def test_dist_factor_and_night_window(over):
    import compat
    import policies as opol
    from fleetrl_b200.policies import night_params
    from fleetrl_b200.tables import build_fleet

    cfg = compat.base_config(**over)
    env = compat.make_reference_env(cfg, start_time="2020-03-02 05:00")
    built = build_fleet(cfg, auto_reset=False)
    c, tb = built.consts, built.tables
    env.reset()
    rng = np.random.default_rng(0)
    for s in range(60):
        t = int(np.searchsorted(built.dates, np.datetime64(env.episode.time)))
        ref = np.asarray(env.get_dist_factor(), dtype=np.float64)                       # fleet_environment.py:782-799
        mine = opol.dist_factor(c, tb, t, np.asarray(env.target_soc, dtype=np.float64))
        np.testing.assert_array_equal(mine, ref, err_msg=f"step {s}")
        env.step(rng.uniform(-1, 1, c.num_evs).astype(np.float32).astype(np.float64))

    # night_charging.py:53-71 on the reference's db
    df = env.db
    lh = df[(df["Location"].shift() == "home") & (df["Location"] == "driving")]
    edt = lh["date"].dt.time.min()
    earliest_dep = edt.hour + edt.minute / 60
    evse, cap = env.load_calculation.evse_max_power, env.ev_config.init_battery_cap
    max_time_needed = env.ev_config.target_soc * cap / env.ev_config.charging_eff / evse
    starting_time = 24 + (earliest_dep - max_time_needed)
    if starting_time > 24:
        starting_time = 23.99
    minutes = np.asarray([0, 15, 30, 45])
    want = (int(math.modf(starting_time)[1]), int(minutes[np.abs(minutes - int(math.modf(starting_time)[0] * 60)).argmin()]),
            int(max_time_needed))
    got = night_params(built)
    assert (got.charging_hour, got.charging_minute, got.max_hours) == want
    assert got.earliest_dep == earliest_dep
