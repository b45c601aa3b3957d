"""Fast synthetic fleet generator: schedules in the reference's CSV schema, plus synthetic price / load / PV series.

The reference generator (fleetrl/utils/schedule/schedule_generator.py:64-691) walks every 15-minute step of the
year with `df.loc[df.date == step, ...] = ...` assignments ("approx. 20 min per EV",
docs/guide/custom_env.rst:151).  This one draws the same per-day random variables from the same distributions
with the same clipping (fleetrl/utils/schedule/schedule_config.py:26-172) but fills whole days at once, so a
50-EV year takes about a second.  It does not reproduce the reference's NumPy legacy RNG draw ORDER (the
reference interleaves one consumption draw per driving step between the daily draws), so schedules are
statistically equivalent, not sample-identical; parity runs therefore feed the SAME generated CSV / tables to
the reference, the oracle and the GPU.

Statistics reproduced
  Delivery  (lmd): weekdays dep N(7,1)h in [3,11], ret N(19,1)h in [12,23], distance N(150,25) km in [20,280];
                   Saturday 9±1.5 / 17±1.5 / 75±25; no Sunday; consumption N(0.213,0.1675) kWh/km in
                   [0.0994,0.453], at most 50 kWh per trip; EVSE 11 kW.
  Utility   (ut) : same shape with its own numbers, 5 % Sunday operation (weekend statistics); EVSE 22 kW.
  Custom         : the delivery pattern with every statistic taken from the env_config's "custom_*" keys.
  Caretaker (ct) : two tours per day around a lunch pause (12:00±15' to 13:30±15' weekdays), operates every day,
                   2 % night emergencies 02:00-04:00 the following night; EVSE 4.7 kW in the file (the env uses 4.6).
Times snap to the 15-minute grid exactly like the reference (int() of the minute fraction, nearest of 0/15/30/45).
"""
import numpy as np
import pandas as pd

_STATS = {
    "lmd": dict(dep_wd=(7, 1), ret_wd=(19, 1), dep_we=(9, 1.5), ret_we=(17, 1.5), dist_wd=(150, 25), dist_we=(75, 25),
                min_dist=20, max_dist=280, cons=(0.213, 0.167463672468669, 0.0994, 0.453), clip=50, min_dep=3,
                max_dep=11, min_ret=12, max_ret=23, power=11, sunday_prob=0.0),
    "ut": dict(dep_wd=(7, 1), ret_wd=(19, 1), dep_we=(9, 2), ret_we=(16, 2), dist_wd=(120, 30), dist_we=(80, 25),
               min_dist=20, max_dist=220, cons=(0.224, 0.167463672468669, 0.0994, 0.453), clip=41, min_dep=3,
               max_dep=11, min_ret=12, max_ret=23, power=22, sunday_prob=0.05),
    "ct": dict(dep_wd=(6, 1), ret_wd=(19, 1), dep_we=(9, 1.5), ret_we=(15, 1.5), pb_wd=(12, 0.25), pe_wd=(13.5, 0.25),
               pb_we=(12, 0.25), pe_we=(13, 0.25), dist_wd=(30, 10), dist_we=(15, 15), min_dist=5, max_dist=50,
               cons=(0.17, 0.167463672468669, 0.0994, 0.453), clip=13.5, clip_pm=10, min_dep=3, max_dep=10,
               min_ret_wd=15, min_ret_we=15, max_ret=23, power=4.7, prob_em=0.02, dist_em=(15, 5), min_em=5),
}


def _custom_stats(env_config):
    """ScheduleType.Custom (schedule_config.py:134-172): the delivery pattern (one trip per working day, a shorter
    Saturday, no Sunday; schedule_generator.py:561-691) with every statistic read from the "custom_*" keys of the
    env_config, the reference's defaults otherwise."""
    g = (env_config or {}).get
    return dict(
        dep_wd=(g("custom_weekday_departure_time_mean", 7), g("custom_weekday_departure_time_std", 1)),
        ret_wd=(g("custom_weekday_return_time_mean", 19), g("custom_weekday_return_time_std", 1)),
        dep_we=(g("custom_weekend_departure_time_mean", 9), g("custom_weekend_departure_time_std", 1.5)),
        ret_we=(g("custom_weekend_return_time_mean", 17), g("custom_weekend_return_time_std", 1.5)),
        dist_wd=(g("custom_weekday_distance_mean", 300), g("custom_weekday_distance_std", 25)),
        dist_we=(g("custom_weekend_distance_mean", 150), g("custom_weekend_distance_std", 25)),
        min_dist=g("custom_minimum_distance", 20), max_dist=g("custom_max_distance", 400),
        cons=(g("custom_consumption_mean", 1.3), g("custom_consumption_std", 0.167463672468669),
              g("custom_minimum_consumption", 0.3994), g("custom_maximum_consumption", 2.5)),
        clip=g("custom_maximum_consumption_per_trip", 500),
        min_dep=g("custom_earliest_hour_of_departure", 3), max_dep=g("custom_latest_hour_of_departure", 11),
        min_ret=g("custom_earliest_hour_of_return", 12), max_ret=g("custom_latest_hour_of_return", 23),
        power=g("custom_ev_charger_power_in_kw", 120), sunday_prob=0.0)


def _snap(time_h, lo=None, hi=None):
    """hour = int(trunc(t)) clipped; minute = nearest of {0,15,30,45} to int(frac*60) (first wins on ties).
    Returns the step-of-day index on the 15-minute grid."""
    hour = np.trunc(time_h).astype(np.int64)
    if lo is not None:
        hour = np.clip(hour, lo, hi)
    frac_min = np.trunc((time_h - np.trunc(time_h)) * 60).astype(np.int64)
    q = np.abs(np.array([0, 15, 30, 45])[None, :] - frac_min[:, None]).argmin(axis=1)
    return hour * 4 + q


def generate_schedule(use_case="lmd", n_evs=1, start="2020-01-01 00:00", end="2020-12-30 23:59", seed=42, env_config=None):
    """-> DataFrame with the reference schedule columns, stacked by vehicle (ID = 0..n_evs-1), 15-minute grid.
    use_case "custom" takes its statistics from env_config's "custom_*" keys (schedule_config.py:134-172)."""
    if use_case == "custom":
        st = _custom_stats(env_config)
    elif use_case in _STATS:
        st = _STATS[use_case]
    else:
        raise TypeError("Company type not found!")
    rng = np.random.default_rng(seed)
    dates = pd.date_range(start=start, end=end, freq="15min")
    if use_case in ("lmd", "custom"):         # schedule_generator.py:75-86, 572-581: skip leading Sundays
        while dates[0].weekday() == 6:
            dates = dates[96:]
    T = len(dates)
    n_days = (T + 95) // 96
    day_wd = np.array([(dates[0] + pd.Timedelta(days=int(d))).weekday() for d in range(n_days)])
    cm, cs, cmin, cmax = st["cons"]
    frames = []
    for ev in range(n_evs):
        dist = np.zeros(n_days * 96)
        cons = np.zeros(n_days * 96)
        driving = np.zeros(n_days * 96, bool)

        def fill(day_idx, a, b, total_distance, clip):
            """steps [a,b) of each listed day become one trip of total_distance km."""
            for d, s0, s1, td in zip(day_idx, a, b, total_distance):
                if s1 <= s0:
                    continue
                n = s1 - s0
                sl = slice(d * 96 + s0, d * 96 + s1)
                rating = np.minimum(np.minimum(np.maximum(rng.normal(cm, cs, n), cmin), cmax), clip / td)
                dist[sl] = td / n
                cons[sl] = (td / n) * rating
                driving[sl] = True

        if use_case in ("lmd", "ut", "custom"):
            wd = np.nonzero(day_wd < 5)[0]
            sat = np.nonzero(day_wd == 5)[0]
            sun = np.nonzero((day_wd == 6) & (rng.random(n_days) > 1 - st["sunday_prob"]))[0] if st["sunday_prob"] else np.array([], int)
            for days, dk, rk, distk in ((wd, "dep_wd", "ret_wd", "dist_wd"), (np.r_[sat, sun], "dep_we", "ret_we", "dist_we")):
                if len(days) == 0:
                    continue
                days = np.sort(days)
                dep = _snap(rng.normal(*st[dk], len(days)), st["min_dep"], st["max_dep"])
                ret = _snap(rng.normal(*st[rk], len(days)), st["min_ret"], st["max_ret"])
                td = np.clip(rng.normal(*st[distk], len(days)), st["min_dist"], st["max_dist"])
                fill(days, dep, ret, td, st["clip"])
        else:  # caretaker
            for days, sfx, min_ret in ((np.nonzero(day_wd < 5)[0], "wd", st["min_ret_wd"]), (np.nonzero(day_wd >= 5)[0], "we", st["min_ret_we"])):
                if len(days) == 0:
                    continue
                dep = _snap(rng.normal(*st["dep_" + sfx], len(days)), st["min_dep"], st["max_dep"])
                pb = _snap(rng.normal(*st["pb_" + sfx], len(days)))
                pe = _snap(rng.normal(*st["pe_" + sfx], len(days)))
                pe = np.where(pe < pb, pb + 1, pe)                      # schedule_generator.py:250-253
                ret = _snap(rng.normal(*st["ret_" + sfx], len(days)), min_ret, st["max_ret"])
                td = np.clip(rng.normal(*st["dist_" + sfx], len(days)), st["min_dist"], st["max_dist"])
                fill(days, dep, pb, td, st["clip"])
                fill(days, pe, ret, td, st["clip_pm"])
            em = np.nonzero(rng.random(n_days) > 1 - st["prob_em"])[0]
            em = em[em + 1 < n_days] + 1                                 # drawn at 23:45, driven 02:00-04:00(+1 step) next night
            if len(em):
                td = np.maximum(rng.normal(*st["dist_em"], len(em)), st["min_em"])
                fill(em, np.full(len(em), 8), np.full(len(em), 17), td, st["clip"])
        dist, cons, driving = dist[:T], cons[:T], driving[:T]
        frames.append(pd.DataFrame({
            "date": dates, "Distance_km": dist, "Consumption_kWh": cons,
            "Location": np.where(driving, "driving", "home"), "ChargingStation": np.where(driving, "none", "home"),
            "ID": ev, "PowerRating_kW": np.where(driving, 0.0, float(st["power"])),
        }))
    return pd.concat(frames, ignore_index=True)


def synthetic_series(start="2020-01-01 00:00", end="2020-12-30 23:59", seed=7, tariff="spot", peak_load_kw=80.0,
                     peak_pv_kw=60.0):
    """Hourly synthetic spot price (EUR/MWh, occasional negative hours), feed-in tariff, building load and PV in the
    reference's column layout.  Shapes only matter for realism; any values are valid inputs of the step."""
    rng = np.random.default_rng(seed)
    hours = pd.date_range(start=pd.Timestamp(start).floor("h"), end=pd.Timestamp(end).floor("h"), freq="h")
    h = hours.hour.values
    doy = hours.dayofyear.values
    season = np.cos((doy - 15) * 2 * np.pi / 365)
    price = 38 + 9 * season + 14 * np.sin((h - 8) * np.pi / 12) + 8 * np.sin((h - 18) * np.pi / 6) + rng.normal(0, 9, len(hours))
    price = np.round(price - 18 * (rng.random(len(hours)) < 0.03) * rng.uniform(1, 4, len(hours)), 2)
    tar = price.copy() if tariff == "spot" else np.full(len(hours), 60.2)
    weekday = hours.weekday.values < 5
    load = peak_load_kw * (0.35 + 0.55 * np.clip(np.sin((h - 6) * np.pi / 13), 0, None) * np.where(weekday, 1.0, 0.45))
    load = np.round(load * rng.uniform(0.92, 1.08, len(hours)), 6)
    sun = np.clip(np.sin((h - 6) * np.pi / 12), 0, None) * (0.55 - 0.4 * season)
    pv = np.round(peak_pv_kw * sun * rng.uniform(0.3, 1.0, len(hours)), 6)
    return (pd.DataFrame({"date": hours, "DELU": price}), pd.DataFrame({"date": hours, "tariff": tar}),
            pd.DataFrame({"date": hours, "load": load}), pd.DataFrame({"date": hours, "pv": pv}))


def write_reference_csvs(path, schedule_name, schedule, price, tariff, building, pv):
    """Write the synthetic inputs in the exact CSV dialects the reference DataLoader parses
    (data_processing.py:47,271-280,307,331,357), so the same files can feed the unmodified reference."""
    import os
    os.makedirs(path, exist_ok=True)
    schedule.to_csv(os.path.join(path, schedule_name))
    p = pd.DataFrame({"date": price["date"]})
    p["Deutschland/Luxemburg [€/MWh] Original resolutions"] = price["DELU"]
    for k in range(16):
        p[f"other_{k}"] = 0.0                                           # the reference drops columns 4:20
    # columns: date, DELU, 16 fillers -> reference drops index 4..19, keeps date/DELU/other_0/other_1
    p.to_csv(os.path.join(path, "synthetic_spot.csv"), sep=";", decimal=",", index=False)
    tariff.to_csv(os.path.join(path, "synthetic_tariff.csv"), sep=";", decimal=",", index=False)
    bl = building.merge(pv, on="date")
    bl.to_csv(os.path.join(path, "synthetic_load.csv"), index=False)
    return dict(schedule_name=schedule_name, price_name="synthetic_spot.csv", tariff_name="synthetic_tariff.csv",
                building_name="synthetic_load.csv", data_path=path)
