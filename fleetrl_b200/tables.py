"""Host-side table builder: the reference DataLoader's semantics as vectorised NumPy, producing the dense tables
that fleet_create() stages in HBM (include/fleetstep.h: FleetTables) plus the FleetConsts scalars.

Follows, and is tested against the live reference `db` in the build container (tests/test_tables_vs_reference.py):
  DataLoader.__init__            fleetrl/utils/data_processing/data_processing.py:21-118   (resample, merge series)
  compute_from_schedule          :120-223   (There, trip grouping, merge_asof backward/forward, SOC_on_return)
  load_prices / load_feed_in / load_building_load / load_pv   :261-370 (CSV dialects, merge_asof backward)
  shape_price_reward             :373-416   (monthly de-trending)
  _date_checker                  :419-431
  FleetEnv.adjust_caretaker_lunch_soc    fleetrl/fleet_env/fleet_environment.py:951-967
  time pickers' candidate ranges         fleetrl/utils/time_picker/*.py
This runs once per environment construction on the host (pandas only parses the CSVs); it is not on the step path.
"""
from dataclasses import dataclass

import numpy as np
import pandas as pd

from . import config as _config
from ._abi import FLEET_DEG_EMPIRICAL, FLEET_DEG_SEI, FleetConsts

SCHEDULE_COLUMNS = ["date", "Distance_km", "Consumption_kWh", "Location", "ChargingStation", "ID", "PowerRating_kW"]


@dataclass
class FleetInputs:
    """Raw inputs in memory (what the reference reads from CSV)."""
    schedule: pd.DataFrame                 # reference schedule schema, stacked by ID, regular grid
    price: pd.DataFrame                    # columns date, DELU  (hourly)
    tariff: pd.DataFrame                   # columns date, tariff
    building: pd.DataFrame = None          # columns date, load
    pv: pd.DataFrame = None                # columns date, pv

    def csv_round_trip(self) -> "FleetInputs":
        """The same inputs as the reference would SEE them after schedule.write_reference_csvs: pandas writes floats with
        repr() and the reference reads them back with read_csv's default (fast, not round-trip exact) float parser, which
        returns a neighbouring float64 for roughly a quarter of the values (1 ulp).  A fleet built from csv_round_trip()
        inputs is therefore bit-identical to the one the unmodified reference builds from the written files
        (tests/test_host_logic.py); the raw in-memory frames give a fleet the reference can never read bit for bit.
        Only the float columns take the text round trip (dates are exact); PV is parsed by Python's float() in the
        reference (decimal="," dialect, data_processing.py:357-360), which is exact."""
        sched = self.schedule.copy()
        for col in ("Distance_km", "Consumption_kWh", "PowerRating_kW"):
            sched[col] = _csv_float_round_trip(sched[col].values, ",", ".")
        price, tariff = self.price.copy(), self.tariff.copy()
        price["DELU"] = _csv_float_round_trip(price["DELU"].values, ";", ",")
        tariff["tariff"] = _csv_float_round_trip(tariff["tariff"].values, ";", ",")
        building = None
        if self.building is not None:
            building = self.building.copy()
            building["load"] = _csv_float_round_trip(building["load"].values, ",", ".")
        return FleetInputs(sched, price, tariff, building, None if self.pv is None else self.pv.copy())


def _csv_float_round_trip(values, sep, decimal):
    import io
    txt = pd.DataFrame({"v": np.asarray(values, np.float64), "pad": "x"}).to_csv(index=False, sep=sep, decimal=decimal)
    return pd.read_csv(io.StringIO(txt), delimiter=sep, decimal=decimal)["v"].to_numpy(np.float64)


def read_inputs(rc: _config.ResolvedConfig) -> FleetInputs:
    """CSV dialects of data_processing.py:47,271-280,307,331,357-360."""
    import os
    cfg = rc.cfg
    path = cfg["data_path"]
    sched = pd.read_csv(os.path.join(path, cfg["schedule_name"]), parse_dates=["date"])
    spot = pd.read_csv(os.path.join(path, cfg["price_name"]), delimiter=";", decimal=",", parse_dates=["date"])
    spot = spot.drop(columns=spot.columns[4:20])
    spot = spot.rename(columns={"Deutschland/Luxemburg [€/MWh] Original resolutions": "DELU"})
    tariff = pd.read_csv(os.path.join(path, cfg["tariff_name"]), delimiter=";", decimal=",", parse_dates=["date"])
    building = pv = None
    if cfg["include_building"]:
        building = pd.read_csv(os.path.join(path, cfg["building_name"]), delimiter=",", parse_dates=["date"])
    if cfg["include_pv"]:
        pv_name = cfg["pv_name"] if cfg["pv_name"] is not None else cfg["building_name"]
        pv = pd.read_csv(os.path.join(path, pv_name), delimiter=",", decimal=",", parse_dates=["date"])
        pv["pv"] = pv["pv"].astype(float)
    return FleetInputs(sched, spot[["date", "DELU"]], tariff[["date", "tariff"]], building, pv)


def _to_ns(a):
    return np.asarray(a, dtype="datetime64[ns]").astype(np.int64)


def _date_checker(df, first_date, last_date):
    """data_processing.py:419-431: shift the series to the schedule's start year, then assert alignment."""
    df = df.copy()
    y_in = pd.Timestamp(df["date"].iloc[0]).year
    y_sched = pd.Timestamp(first_date).year
    if y_in != y_sched:
        df["date"] = df["date"] + pd.DateOffset(years=y_sched - y_in)
    assert pd.Timestamp(df["date"].iloc[0]) == pd.Timestamp(first_date), "Invalid start time."
    assert pd.Timestamp(df["date"].iloc[-1]).year == pd.Timestamp(last_date).year, "Invalid end year."
    return df


def _merge_backward(grid_ns, df, col, first_date, last_date):
    """pd.merge_asof(date_range, series.sort_values('date'), direction='backward') as searchsorted."""
    df = _date_checker(df, first_date, last_date).sort_values("date")  # default (unstable) sort, like the reference
    d = _to_ns(df["date"].values)
    v = df[col].values.astype(np.float64)
    idx = np.searchsorted(d, grid_ns, side="right") - 1
    out = np.where(idx >= 0, v[np.clip(idx, 0, len(v) - 1)], np.nan)
    return out


def _resample_schedule(s: pd.DataFrame, minutes: int) -> pd.DataFrame:
    """groupby('ID').resample(freq).agg(...) of data_processing.py:58-61 for a grid coarser than the data."""
    d = s["date"].values.astype("datetime64[ns]")
    step_in = int((d[1] - d[0]) / np.timedelta64(1, "m"))
    if step_in == minutes:
        return s.reset_index(drop=True)
    if minutes % step_in != 0 or minutes < step_in:
        raise ValueError("up-sampling / non-integer resampling of the schedule is not supported (neither does the reference)")
    bucket = s["date"].dt.floor(f"{minutes}min")
    g = s.groupby([s["ID"].values, bucket.values], sort=True)
    out = g.agg({'Location': 'first', 'ID': 'first', 'Consumption_kWh': 'sum', 'ChargingStation': 'first',
                 'PowerRating_kW': 'mean', 'Distance_km': 'sum', 'date': 'first'})
    return out.reset_index(drop=True)


def _kahan_segment_sum(v, starts, ends):
    """Per-segment Kahan-compensated sums, the algorithm pandas' groupby().sum() uses (pandas/_libs/groupby.pyx
    group_sum), vectorised over segments so that the result is bit-identical to the reference's trip consumption."""
    lens = ends - starts
    total = np.zeros(len(starts))
    comp = np.zeros(len(starts))
    for k in range(int(lens.max()) if len(lens) else 0):
        m = lens > k
        val = v[starts[m] + k]
        y = val - comp[m]
        t = total[m] + y
        comp[m] = t - total[m] - y
        total[m] = t
    return total


def compute_from_schedule(s: pd.DataFrame, minutes: int, target_soc: float, init_battery_cap: float):
    """compute_from_schedule (data_processing.py:120-223) on the stacked (ID, date)-sorted frame, vectorised.

    Returns there[N,T] u8, time_left[N,T] f64, soc_on_return[N,T] f64, consumption[N,T] f64, dates[T].
    The trip grouping is done on the STACKED frame exactly like the reference (Change/Group via shift over the
    whole frame, :133-136), including its attribution of a group to the first ID it touches.
    """
    ids = s["ID"].values.astype(np.int64)
    N = int(ids.max()) + 1
    if len(s) % N != 0:
        raise ValueError("every vehicle must cover the same date grid")
    T = len(s) // N
    date_ns = _to_ns(s["date"].values)
    if not (np.array_equal(ids, np.repeat(np.arange(N), T)) and np.array_equal(date_ns.reshape(N, T), np.tile(date_ns[:T], (N, 1)))):
        raise ValueError("schedule must be stacked by ID with one identical, sorted date grid per vehicle")
    cs = s["ChargingStation"].values.astype(str)
    there = (s["PowerRating_kW"].values != 0)
    cons_step = s["Consumption_kWh"].values.astype(np.float64)
    none = cs == "none"
    change = np.ones(len(s), bool)
    change[1:] = cs[1:] != cs[:-1]                                                   # :133
    group = np.cumsum(change)                                                        # :136
    # groups restricted to rows with ChargingStation == "none" (:145-167)
    gi = np.nonzero(none)[0]
    consumption = np.zeros(len(s))
    time_left = np.zeros(len(s))
    if len(gi):
        g = group[gi]
        starts = np.nonzero(np.r_[True, g[1:] != g[:-1]])[0]
        ends = np.r_[starts[1:], len(gi)]
        trip_cons = _kahan_segment_sum(cons_step[gi], starts, ends)                  # groupby.sum (pandas: Kahan)
        first_row, last_row = gi[starts], gi[ends - 1]
        trip_id = ids[first_row]                                                     # ["ID"].first()
        ret_ns = date_ns[last_row] + np.int64(minutes) * 60_000_000_000              # :163
        dep_ns = date_ns[first_row]
        # merge_asof backward by ID: last return event with return_date <= date (:176-181)
        cons_nt = np.full(len(s), np.nan)
        dep_nt = np.full(len(s), np.nan)
        for n in range(N):
            sel = np.nonzero(trip_id == n)[0]
            rows = slice(n * T, (n + 1) * T)
            if len(sel) == 0:
                continue
            order = np.argsort(ret_ns[sel], kind="stable")
            r_d, r_c = ret_ns[sel][order], trip_cons[sel][order]
            k = np.searchsorted(r_d, date_ns[rows], side="right") - 1
            cons_nt[rows] = np.where(k >= 0, r_c[np.clip(k, 0, len(r_c) - 1)], np.nan)
            order = np.argsort(dep_ns[sel], kind="stable")
            d_d = dep_ns[sel][order]
            k = np.searchsorted(d_d, date_ns[rows], side="left")                     # forward (:196-201)
            dep_nt[rows] = np.where(k < len(d_d), d_d[np.clip(k, 0, len(d_d) - 1)].astype(np.float64), np.nan)
        cons_nt[~there] = 0                                                          # :187
        consumption = np.nan_to_num(cons_nt, nan=0.0)                                # :193
        tl = (dep_nt - date_ns.astype(np.float64)) / 1e9 / 3600                      # :206 total_seconds()/3600
        tl[~there] = 0                                                               # :208
        time_left = np.nan_to_num(tl, nan=0.0)                                       # :210 (pandas-2.2 semantics)
    sr = target_soc - consumption / init_battery_cap                                 # :221-222
    sr[~there] = 0                                                                   # :223
    shape = (N, T)
    return (there.astype(np.uint8).reshape(shape), time_left.reshape(shape), sr.reshape(shape),
            consumption.reshape(shape), s["date"].values[:T].astype("datetime64[ns]"))


def shape_reward_curve(values, dates, offset, factor):
    """shape_price_reward (data_processing.py:387-414): (x + offset) * factor, each calendar month shifted so that
    its mean equals the mean of the whole series."""
    x = pd.Series((values + offset) * factor if offset is not None else values * factor)
    total_avg = x.mean()
    idx = pd.DatetimeIndex(dates)
    key = idx.year.values * 12 + idx.month.values
    out = np.empty(len(x))
    starts = np.nonzero(np.r_[True, key[1:] != key[:-1]])[0]
    ends = np.r_[starts[1:], len(x)]
    for a, b in zip(starts, ends):
        chunk = x.iloc[a:b]
        out[a:b] = (chunk - chunk.mean() + total_avg).values
    return out


def calendar_tables(dates):
    idx = pd.DatetimeIndex(dates)
    month, wd, hour, minute = idx.month.values, idx.weekday.values, idx.hour.values, idx.minute.values
    # evaluate the reference's scalar expressions (observer_bl_pv.py:100-107) once per distinct value
    ms = {m: (np.sin(2 * np.pi * m / 12), np.cos(2 * np.pi * m / 12)) for m in range(1, 13)}
    ws = {w: (np.sin(2 * np.pi * w / 7), np.cos(2 * np.pi * w / 7)) for w in range(7)}
    hs = {h: (np.sin(2 * np.pi * h / 24), np.cos(2 * np.pi * h / 24)) for h in range(24)}
    cal = np.empty((len(idx), 6))
    cal[:, 0] = [ms[int(m)][0] for m in month]; cal[:, 1] = [ms[int(m)][1] for m in month]
    cal[:, 2] = [ws[int(w)][0] for w in wd];    cal[:, 3] = [ws[int(w)][1] for w in wd]
    cal[:, 4] = [hs[int(h)][0] for h in hour];  cal[:, 5] = [hs[int(h)][1] for h in hour]
    return cal, hour.astype(np.uint8), minute.astype(np.uint8)


@dataclass
class BuiltFleet:
    consts: FleetConsts
    tables: dict
    dates: np.ndarray            # [T] datetime64[ns]
    company: _config.Company
    rc: _config.ResolvedConfig
    start_ranges: dict           # time picker name -> (lo, hi) inclusive index range, or a fixed index for "static"
    consumption: np.ndarray = None


CACHE_VERSION = 1


def save_fleet(built: "BuiltFleet", path: str):
    """Binary cache of a built fleet: the flattened tables, the scalar half of FleetConsts and what the
    host wrappers need (dates, start ranges, company).  The reference rebuilds its `db` from the CSV files at every
    construction (~10 s of pandas at N = 50); loading this file takes milliseconds and is independent of pandas."""
    import dataclasses
    import json
    arrays = {f"t_{k}": v for k, v in built.tables.items() if v is not None}
    meta = dict(version=CACHE_VERSION, consts=built.consts.to_dict(), company=dataclasses.asdict(built.company),
                start_ranges={k: list(v) for k, v in built.start_ranges.items()},
                absent=[k for k, v in built.tables.items() if v is None], cfg=built.rc.cfg)
    np.savez_compressed(path, dates=np.asarray(built.dates).astype("datetime64[ns]").astype(np.int64),
                        consumption=built.consumption if built.consumption is not None else np.zeros(0),
                        meta=np.frombuffer(json.dumps(meta, default=str).encode(), dtype=np.uint8), **arrays)


def load_fleet(path: str) -> "BuiltFleet":
    """Inverse of save_fleet: tables and constants are restored bit for bit."""
    import json
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    meta = json.loads(bytes(z["meta"]).decode())
    if meta.get("version") != CACHE_VERSION:
        raise ValueError(f"{path}: fleet cache version {meta.get('version')} != {CACHE_VERSION}")
    tables = {k[2:]: z[k] for k in z.files if k.startswith("t_")}
    for k in meta["absent"]:
        tables[k] = None
    cfg = meta["cfg"]
    rc = _config.resolve(cfg) if all(k in cfg for k in _config.MANDATORY_KEYS) else None
    cons = z["consumption"]
    return BuiltFleet(consts=FleetConsts.from_dict(meta["consts"]), tables=tables,
                      dates=z["dates"].astype("datetime64[ns]"), company=_config.Company(**meta["company"]), rc=rc,
                      start_ranges={k: tuple(v) for k, v in meta["start_ranges"].items()},
                      consumption=cons if cons.size else None)


def start_index_ranges(dates, time_conf, static_start="01/02/2021 19:00"):
    """Candidate start indices of the three time pickers (time_picker/*.py)."""
    idx = pd.DatetimeIndex(dates)
    T = len(idx)
    step = np.timedelta64(int(time_conf.minutes), "m")
    last = idx[-1]
    def pos(ts):
        return int(np.clip((pd.Timestamp(ts) - idx[0]) // pd.Timedelta(step), 0, T - 1))
    rnd_hi = pos(last - pd.Timedelta(days=time_conf.end_cutoff))                     # random_time_picker.py:25-28
    ev_lo = rnd_hi                                                                   # eval_time_picker.py:33-36
    ev_hi = pos(last - pd.Timedelta(hours=2 * time_conf.episode_length))
    st = pd.to_datetime(static_start)                                                # static_time_picker.py:20-29
    if st.year < idx[0].year or st.year > idx[-1].year:
        st = st + pd.DateOffset(years=idx[0].year - st.year)
    return {"random": (0, rnd_hi), "eval": (ev_lo, max(ev_lo, ev_hi)), "static": (pos(st), pos(st))}


def build_fleet(env_config, inputs: FleetInputs = None, *, auto_reset=True, carry_degradation_state=True,
                seed=None, time_picker=None) -> BuiltFleet:
    """env_config (dict or JSON path, reference keys) -> FleetConsts + canonical tables."""
    rc = _config.resolve(env_config)
    cfg, ev, sc, tc = rc.cfg, rc.ev, rc.score, rc.time
    if cfg["gen_schedule"]:
        # FleetEnv.auto_gen (fleet_environment.py:181,969-992): generate gen_n_evs schedules over gen_start_date ..
        # gen_end_date, save them as gen_name next to the other inputs and use that file (statistically equivalent fast
        # generator, fleetrl_b200/schedule.py; with in-memory inputs the generated frame just replaces inputs.schedule)
        import os
        from .schedule import generate_schedule
        sched_gen = generate_schedule(rc.use_case, int(cfg["gen_n_evs"]), cfg["gen_start_date"], cfg["gen_end_date"],
                                      seed=int(cfg["seed"]), env_config=cfg)
        if inputs is None:
            name = cfg["gen_name"] if str(cfg["gen_name"]).endswith(".csv") else str(cfg["gen_name"]) + ".csv"
            sched_gen.to_csv(os.path.join(cfg["data_path"], name))
            cfg["schedule_name"] = name
        else:
            inputs = FleetInputs(sched_gen, inputs.price, inputs.tariff, inputs.building, inputs.pv)
    if inputs is None:
        inputs = read_inputs(rc)
    sched = inputs.schedule.copy()
    if "date" in sched.columns:
        sched["date"] = pd.to_datetime(sched["date"])
    sched = _resample_schedule(sched, tc.minutes)
    there, time_left, sr, consumption, dates = compute_from_schedule(sched, tc.minutes, ev.target_soc, ev.init_battery_cap)
    N, T = there.shape
    # date_range(start=min, end=max, freq) — must equal the schedule grid (data_processing.py:75-77)
    grid = pd.date_range(start=dates[0], end=dates[-1], freq=f"{tc.minutes}min")
    if len(grid) != T:
        raise ValueError("schedule dates are not a regular grid at the configured frequency")
    grid_ns = _to_ns(grid.values)
    first, last = grid[0], grid[-1]
    delu = _merge_backward(grid_ns, inputs.price, "DELU", first, last)
    tariff = _merge_backward(grid_ns, inputs.tariff, "tariff", first, last)
    load = _merge_backward(grid_ns, inputs.building, "load", first, last) if cfg["include_building"] else None
    pv = _merge_backward(grid_ns, inputs.pv, "pv", first, last) if cfg["include_pv"] else None

    if rc.use_case == "ct":                                                          # fleet_environment.py:951-967
        hour = pd.DatetimeIndex(dates).hour.values
        sel = ((hour >= 0) & (hour <= 10)) | ((hour >= 15) & (hour <= 23))
        sr[:, sel] = ev.target_soc_lunch - consumption[:, sel] / ev.init_battery_cap
        sr[there == 0] = 0

    max_load = float(np.max(load)) if cfg["include_building"] else 0                  # :265-268
    company = _config.company_for(rc.use_case, cfg, max_load, N)
    prc = shape_reward_curve(delu, dates, ev.fixed_markup, ev.variable_multiplier)
    trc = shape_reward_curve(tariff, dates, None, 1 - ev.feed_in_deduction)
    cal, hour_u8, minute_u8 = calendar_tables(dates)

    ranges = start_index_ranges(dates, tc)
    tp = time_picker or cfg["time_picker"]
    if tp not in ranges:
        raise TypeError("Time picker type not recognised")
    lo, hi = ranges[tp]
    sph = int(1 / tc.dt)                                                             # fleet_environment.py:456
    consts = FleetConsts.from_dict(dict(
        num_evs=N, table_len=T, steps_per_hour=sph, episode_steps=int(tc.episode_length * sph),
        price_lookahead=tc.price_lookahead, bl_pv_lookahead=tc.bl_pv_lookahead,
        include_price=int(cfg["include_price"]), include_building=int(cfg["include_building"]),
        include_pv=int(cfg["include_pv"]), aux=int(cfg["aux"]), normalize=int(cfg["normalize_in_env"]),
        is_caretaker=int(rc.use_case == "ct"), calc_degradation=int(cfg["calculate_degradation"]),
        deg_mode=FLEET_DEG_EMPIRICAL if cfg["deg_emp"] else FLEET_DEG_SEI,
        carry_degradation_state=int(carry_degradation_state), auto_reset=int(auto_reset),
        start_lo=lo, start_hi=hi, seed=int(cfg["seed"] if seed is None and cfg["seed"] is not None else (seed or 0)),
        dt=tc.dt, init_battery_cap=ev.init_battery_cap, obc_max_power=ev.obc_max_power, charging_eff=ev.charging_eff,
        discharging_eff=ev.discharging_eff, def_soc=ev.def_soc, temperature=ev.temperature, target_soc=ev.target_soc,
        target_soc_lunch=ev.target_soc_lunch, min_laxity=ev.min_laxity, fixed_markup=ev.fixed_markup,
        variable_multiplier=ev.variable_multiplier, feed_in_deduction=ev.feed_in_deduction,
        evse_max_power=company.evse_max_power, grid_connection=company.grid_connection, lc_batt_cap=company.batt_cap,
        price_multiplier=sc.price_multiplier, fully_charged_reward=sc.fully_charged_reward,
        penalty_invalid_action=sc.penalty_invalid_action, penalty_overcharging=sc.penalty_overcharging,
        penalty_overloading=sc.penalty_overloading, clip_overcharging=sc.clip_overcharging,
        init_soh=cfg["init_soh"], soc_eps=0.005,
        max_time_left=float(np.max(time_left)),                                      # oracle_normalization.py:34-47
        min_price=(float(np.min(delu)) + ev.fixed_markup) * ev.variable_multiplier,
        max_price=(float(np.max(delu)) + ev.fixed_markup) * ev.variable_multiplier,
        min_tariff=float(np.min(tariff)) * (1 - ev.feed_in_deduction),
        max_tariff=float(np.max(tariff)) * (1 - ev.feed_in_deduction),
        max_building=max_load if cfg["include_building"] else 0.0,
        max_pv=float(np.max(pv)) if cfg["include_pv"] else 0.0,
    ))
    tables = dict(there=there, time_left=time_left, soc_on_return=sr, delu=delu, tariff=tariff, load=load, pv=pv,
                  price_reward_curve=prc, tariff_reward_curve=trc, cal_sincos=cal, hour=hour_u8, minute=minute_u8)
    return BuiltFleet(consts=consts, tables=tables, dates=dates, company=company, rc=rc, start_ranges=ranges,
                      consumption=consumption)
