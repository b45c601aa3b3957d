"""The fast schedule generator (fleetrl_b200/schedule.py) against the live reference ScheduleConfig / ScheduleGenerator
(build container only).  The generator is statistically equivalent, not RNG-identical (schedule.py docstring), so:
  * every statistic it draws from is compared VALUE BY VALUE with the unmodified ScheduleConfig objects,
  * a three-week schedule from the unmodified ScheduleGenerator and a long one from ours are compared structurally
    (schema, value sets, trip structure per weekday type, clipping bounds) and by two-sample Kolmogorov-Smirnov tests on
    departure step, return step, trip distance and per-step consumption rating."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.reference
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "refshim"))

CUSTOM = {"custom_weekday_distance_mean": 220, "custom_consumption_mean": 1.1, "custom_latest_hour_of_return": 22,
          "custom_ev_charger_power_in_kw": 150, "custom_maximum_consumption_per_trip": 420}


def _ref_config(use_case):
    import compat
    compat.install()
    from fleetrl.utils.schedule.schedule_config import ScheduleConfig, ScheduleType
    st = {"lmd": ScheduleType.Delivery, "ct": ScheduleType.Caretaker, "ut": ScheduleType.Utility, "custom": ScheduleType.Custom}[use_case]
    return ScheduleConfig(schedule_type=st, env_config=dict(CUSTOM)), st


@pytest.mark.parametrize("use_case", ["lmd", "ut", "ct", "custom"])
def test_statistics_equal_reference_schedule_config(use_case):
    from fleetrl_b200.schedule import _STATS, _custom_stats
    sc, _ = _ref_config(use_case)
    st = _custom_stats(CUSTOM) if use_case == "custom" else _STATS[use_case]
    pairs = {"dep_wd": ("dep_mean_wd", "dep_dev_wd"), "ret_wd": ("ret_mean_wd", "ret_dev_wd"), "dep_we": ("dep_mean_we", "dep_dev_we"),
             "ret_we": ("ret_mean_we", "ret_dev_we"), "dist_wd": ("avg_distance_wd", "dev_distance_wd"),
             "dist_we": ("avg_distance_we", "dev_distance_we")}
    if use_case == "ct":
        pairs.update({"pb_wd": ("pause_beg_mean_wd", "pause_beg_dev_wd"), "pe_wd": ("pause_end_mean_wd", "pause_end_dev_wd"),
                      "pb_we": ("pause_beg_mean_we", "pause_beg_dev_we"), "pe_we": ("pause_end_mean_we", "pause_end_dev_we")})
    for k, (m, d) in pairs.items():
        assert st[k] == (getattr(sc, m), getattr(sc, d)), k
    assert st["cons"] == (sc.consumption_mean, sc.consumption_std, sc.consumption_min, sc.consumption_max)
    assert st["clip"] == sc.total_cons_clip and st["power"] == sc.charging_power
    assert (st["min_dist"], st["max_dist"], st["min_dep"], st["max_dep"], st["max_ret"]) == \
           (sc.min_distance, sc.max_distance, sc.min_dep, sc.max_dep, sc.max_return_hour)
    if use_case == "ct":
        assert (st["min_ret_wd"], st["min_ret_we"], st["clip_pm"], st["prob_em"], st["dist_em"], st["min_em"]) == \
               (sc.min_return_wd, sc.min_return_we, sc.total_cons_clip_afternoon, sc.prob_emergency,
                (sc.avg_distance_em, sc.dev_distance_em), sc.min_em_distance)
    else:
        assert st["min_ret"] == sc.min_return


def _trips(df):
    """one record per trip: weekday, departure step of day, return step, distance, per-step ratings"""
    df = df.reset_index(drop=True)
    drv = (df["Location"] == "driving").to_numpy()
    edges = np.flatnonzero(np.diff(np.r_[0, drv.astype(int), 0]))
    out = []
    for a, b in zip(edges[::2], edges[1::2]):
        ts = pd.Timestamp(df["date"].iloc[a])
        dist = df["Distance_km"].iloc[a:b].to_numpy()
        out.append(dict(wd=ts.weekday(), dep=ts.hour * 4 + ts.minute // 15, n=b - a, dist=dist.sum(),
                        rating=df["Consumption_kWh"].iloc[a:b].to_numpy() / dist))
    return out


@pytest.mark.parametrize("use_case", ["lmd", "ut"])
def test_distributions_match_reference_generator(use_case):
    from scipy.stats import ks_2samp
    import compat
    from fleetrl_b200.schedule import generate_schedule
    _, st = _ref_config(use_case)
    from fleetrl.utils.schedule.schedule_generator import ScheduleGenerator
    cfg = compat.base_config(use_case=use_case, gen_start_date="2020-01-06 00:00", gen_end_date="2020-01-26 23:59", seed=7,
                             freq="15T")
    ref = ScheduleGenerator(env_config=cfg, schedule_type=st, vehicle_id="0").generate_schedule()
    mine = generate_schedule(use_case, 1, start="2020-01-06 00:00", end="2020-12-27 23:59", seed=3)
    assert list(mine.columns) == ["date", "Distance_km", "Consumption_kWh", "Location", "ChargingStation", "ID", "PowerRating_kW"]
    assert set(ref.columns) == set(mine.columns)
    for col in ("Location", "ChargingStation"):
        assert set(ref[col]) == set(mine[col])
    assert set(np.unique(ref["PowerRating_kW"])) == set(np.unique(mine["PowerRating_kW"]))
    tr, tm = _trips(ref), _trips(mine)
    # same trip structure: one trip per working day, none on the reference's Sundays
    assert len([t for t in tr if t["wd"] < 5]) == 15 and len([t for t in tm if t["wd"] < 5]) == 51 * 5 - 1 + 1
    if use_case == "lmd":
        assert not [t for t in tr if t["wd"] == 6] and not [t for t in tm if t["wd"] == 6]
    for key in ("dep", "n", "dist"):
        a = np.array([t[key] for t in tr if t["wd"] < 5], float)
        b = np.array([t[key] for t in tm if t["wd"] < 5], float)
        assert ks_2samp(a, b).pvalue > 1e-3, (key, a.mean(), b.mean())
    ra = np.concatenate([t["rating"] for t in tr])
    rb = np.concatenate([t["rating"] for t in tm])
    assert ks_2samp(ra, rb).pvalue > 1e-3, (ra.mean(), rb.mean())
    assert abs(ra.min() - rb.min()) < 1e-12 and ra.max() <= rb.max() + 1e-12          # same floor; same ceiling family
