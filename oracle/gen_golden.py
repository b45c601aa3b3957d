"""Generate golden trajectories by running the UNMODIFIED reference FleetEnv (/root/reference) in THIS container.

TEST INFRASTRUCTURE.  Run once here (`python oracle/gen_golden.py`); the resulting small .npz files under
tests/golden/ are committed and travel to the GPU box, the reference does not.

Each fixture holds, for one scenario:
  consts   : every scalar the path needs, read off the live reference objects after FleetEnv.__init__
  tables   : the reference's `db` columns, windowed to the rows the episode(s) can touch (+ look-ahead)
  start    : start index inside the window per episode, actions (float32) per step
  traj     : what the reference returned / held after reset and after every step:
             obs (float32), reward, cashflow (Episode.current_charging_expense), done, soc, hours_left, soc_deg,
             soh, target_soc, and after each daily degradation call rainflow_length / fd_cyc / l.

Actions are float32 values handed to the reference as float64 arrays of the same values: the reference is pinned
to NumPy 1.26 where `python_float * np.float32 -> float64`; under this container's NumPy 2 (NEP 50) a float32
action array would silently demote parts of EvCharger.charge to float32.  Widening first reproduces the pinned
behaviour exactly.

Multi-EV schedules: only 1-EV schedules ship (inputs/2_*.csv are missing blobs), so N-EV fleets are built by
stacking the shipped 1-EV schedule N times, vehicle k rolled by k*7 days (keeps the weekday structure), ID=k,
written in the reference CSV schema to a scratch data_path next to symlinks of the price/load files
(SURVEY App. C-5).
"""
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, ROOT)
import compat  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF_INPUTS = os.path.join(compat.REFERENCE_ROOT, "inputs")


def make_data_path(schedule_src, n_evs, name):
    """Scratch input dir with an N-EV stacked schedule + symlinks to every other shipped input."""
    d = tempfile.mkdtemp(prefix="fleet_golden_")
    for f in os.listdir(REF_INPUTS):
        os.symlink(os.path.join(REF_INPUTS, f), os.path.join(d, f))
    df = pd.read_csv(os.path.join(REF_INPUTS, schedule_src), index_col=0)
    steps_per_week = 7 * 96
    parts = []
    cols = [c for c in df.columns if c not in ("date", "ID")]
    for k in range(n_evs):
        p = df.copy()
        for c in cols:
            p[c] = np.roll(df[c].values, k * steps_per_week)
        p["ID"] = k
        parts.append(p)
    pd.concat(parts).reset_index(drop=True).to_csv(os.path.join(d, name))
    return d


def extract_consts(env, deg_mode, carry=True):
    ec, sc, tc, lc = env.ev_config, env.score_config, env.time_conf, env.load_calculation
    db = env.db
    N = int(env.num_cars)
    T = len(db) // N
    dt = tc.minutes / 60
    c = dict(
        num_evs=N, table_len=T, steps_per_hour=int(1 / dt), episode_steps=int(tc.episode_length * int(1 / dt)),
        price_lookahead=tc.price_lookahead, bl_pv_lookahead=tc.bl_pv_lookahead,
        include_price=int(env.include_price), include_building=int(env.include_building_load),
        include_pv=int(env.include_pv), aux=int(env.aux_flag), normalize=int(env.normalize_in_env),
        is_caretaker=int(env.company.name == "Caretaker"), calc_degradation=int(env.calc_deg), deg_mode=deg_mode,
        carry_degradation_state=int(carry), auto_reset=0, start_lo=0, start_hi=0, seed=0,
        dt=dt, init_battery_cap=ec.init_battery_cap, obc_max_power=ec.obc_max_power, charging_eff=ec.charging_eff,
        discharging_eff=ec.discharging_eff, def_soc=ec.def_soc, temperature=ec.temperature,
        target_soc=ec.target_soc, target_soc_lunch=ec.target_soc_lunch, min_laxity=ec.min_laxity,
        fixed_markup=ec.fixed_markup, variable_multiplier=ec.variable_multiplier,
        feed_in_deduction=ec.feed_in_deduction,
        evse_max_power=lc.evse_max_power, grid_connection=lc.grid_connection, lc_batt_cap=lc.batt_cap,
        price_multiplier=sc.price_multiplier, fully_charged_reward=sc.fully_charged_reward,
        penalty_invalid_action=sc.penalty_invalid_action, penalty_overcharging=sc.penalty_overcharging,
        penalty_overloading=sc.penalty_overloading, clip_overcharging=sc.clip_overcharging,
        init_soh=env.initial_soh, soc_eps=env.eps,
        # OracleNormalization scales, oracle_normalization.py:34-47 (same expressions, evaluated on the live db)
        max_time_left=max(db["time_left"]),
        min_price=(min(db["DELU"]) + ec.fixed_markup) * ec.variable_multiplier,
        max_price=(max(db["DELU"]) + ec.fixed_markup) * ec.variable_multiplier,
        min_tariff=(min(db["tariff"])) * (1 - ec.feed_in_deduction),
        max_tariff=(max(db["tariff"])) * (1 - ec.feed_in_deduction),
        max_building=max(db["load"]) if env.include_building_load else 0.0,
        max_pv=max(db["pv"]) if env.include_pv else 0.0,
    )
    return {k: (float(v) if isinstance(v, (float, np.floating)) else int(v)) for k, v in c.items()}


def extract_tables(env, w0, w1):
    db = env.db
    N = int(env.num_cars)
    T = len(db) // N
    dates = pd.DatetimeIndex(db["date"].values[:T])

    def per_ev(col, dtype):
        return db[col].values.reshape(N, T)[:, w0:w1].astype(dtype)

    def per_t(col):
        return db[col].values[:T][w0:w1].astype(np.float64) if col in db.columns else None

    d = dates[w0:w1]
    # observer_bl_pv.py:100-107 — evaluated with the same Python expressions on the same Timestamp fields
    cal = np.array([[np.sin(2 * np.pi * ts.month / 12), np.cos(2 * np.pi * ts.month / 12),
                     np.sin(2 * np.pi * ts.weekday() / 7), np.cos(2 * np.pi * ts.weekday() / 7),
                     np.sin(2 * np.pi * ts.hour / 24), np.cos(2 * np.pi * ts.hour / 24)] for ts in d])
    tb = dict(there=per_ev("There", np.uint8), time_left=per_ev("time_left", np.float64),
              soc_on_return=per_ev("SOC_on_return", np.float64),
              delu=per_t("DELU"), tariff=per_t("tariff"), load=per_t("load"), pv=per_t("pv"),
              price_reward_curve=per_t("price_reward_curve"), tariff_reward_curve=per_t("tariff_reward_curve"),
              cal_sincos=cal, hour=np.array(d.hour, np.uint8), minute=np.array(d.minute, np.uint8))
    return tb, dates


def make_generated_data_path(use_case, n_evs, name, gen_seed):
    """Scratch input dir holding a fleet from the PRODUCT's generator (fleetrl_b200.schedule) written in the reference's
    CSV dialects by write_reference_csvs: the BASELINE.json fleet shapes (50-EV lmd / ut, 20-EV ct), which the shipped
    1-EV schedules cannot provide."""
    from fleetrl_b200.schedule import generate_schedule, synthetic_series, write_reference_csvs
    d = tempfile.mkdtemp(prefix="fleet_golden_gen_")
    sched = generate_schedule(use_case, n_evs, seed=gen_seed)
    price, tariff, load, pv = synthetic_series(seed=gen_seed + 1)
    return write_reference_csvs(d, f"{n_evs}_{name}.csv", sched, price, tariff, load, pv)


def run_case(name, cfg_over, starts, n_steps_per_ep, action_fn, schedule=None, n_evs=1, linear=False, seed=0,
             pre_hook=None, generated=None, log=False):
    """starts: list of start-time strings, one per episode run back to back on the SAME env object.
    generated: seed of a fleet from the product's own generator (written to CSV, read by the reference);
    log: run with log_data=True and keep every DataLogger column (data_logger.py:55-68)."""
    cfg = compat.base_config(**cfg_over)
    if log:
        cfg["log_data"] = True
    if generated is not None:
        cfg.update(make_generated_data_path(cfg["use_case"], n_evs, name, generated))
    elif n_evs > 1 or schedule is not None:
        src = schedule or cfg["schedule_name"]
        sched_name = f"{n_evs}_{name}.csv"
        cfg["data_path"] = make_data_path(src, n_evs, sched_name)
        cfg["schedule_name"] = sched_name
    if linear:
        cfg["deg_emp"] = True
    env = compat.make_reference_env(cfg, start_time=starts[0], linear_degradation_patch=linear)
    if pre_hook:
        pre_hook(env)
    from fleetrl.utils.time_picker.static_time_picker import StaticTimePicker

    N = int(env.num_cars)
    T = len(env.db) // N
    dates = pd.DatetimeIndex(env.db["date"].values[:T])
    sph = int(1 / (env.time_conf.minutes / 60))
    t0s = [int(dates.searchsorted(pd.Timestamp(s))) for s in starts]
    w0 = max(0, min(t0s) - 2)
    w1 = min(T, max(t0s) + n_steps_per_ep + (env.time_conf.price_lookahead + 3) * sph + 2)
    consts = extract_consts(env, deg_mode=1 if linear else 0)
    tables, _ = extract_tables(env, w0, w1)
    consts["table_len"] = w1 - w0
    rng = np.random.default_rng(seed)

    rec = {k: [] for k in "obs reward cashflow done soc hours_left soc_deg soh target_soc rf_len fd_cyc life".split()}
    actions_all, ep_start_rows = [], []

    def snap(obs):
        ep = env.episode
        rec["obs"].append(np.asarray(obs, np.float32).copy())
        rec["soc"].append(np.array(ep.soc, np.float64))
        rec["hours_left"].append(np.array(ep.hours_left, np.float64))
        rec["soc_deg"].append(np.array(ep.soc_deg, np.float64))
        rec["soh"].append(np.array(ep.soh, np.float64))
        rec["target_soc"].append(np.array(env.target_soc, np.float64))
        if not linear:
            rec["rf_len"].append(np.array(env.sei_deg.rainflow_length, np.float64))
            rec["fd_cyc"].append(np.array(env.sei_deg.fd_cyc, np.float64))
            rec["life"].append(np.array(env.sei_deg.l, np.float64))

    for ep_i, s in enumerate(starts):
        env.time_picker = StaticTimePicker(start_time=s)
        obs, _ = env.reset()
        ep_start_rows.append(len(rec["obs"]))
        snap(obs)
        for k in range(n_steps_per_ep):
            a32 = np.asarray(action_fn(rng, k, N, env), np.float32)
            actions_all.append(a32)
            obs, r, done, trunc, info = env.step(a32.astype(np.float64))
            rec["reward"].append(float(r))
            rec["cashflow"].append(float(env.episode.current_charging_expense))
            rec["done"].append(bool(done))
            snap(obs)

    out = {f"tb_{k}": v for k, v in tables.items() if v is not None}
    out.update({f"tr_{k}": np.array(v) for k, v in rec.items() if len(v)})
    out["actions"] = np.array(actions_all, np.float32)
    out["start_idx"] = np.array([t - w0 for t in t0s], np.int32)
    out["ep_start_rows"] = np.array(ep_start_rows, np.int32)
    out["n_steps_per_ep"] = np.int32(n_steps_per_ep)
    if log:
        lg = env.data_logger.log
        T0 = dates[w0]
        step_td = pd.Timedelta(minutes=int(env.time_conf.minutes))
        out["log_episode"] = lg["Episode"].to_numpy(np.int64)
        out["log_time_idx"] = np.array([int((pd.Timestamp(t) - T0) / step_td) for t in lg["Time"]], np.int64)
        out["log_obs"] = np.stack([np.asarray(o, np.float32) for o in lg["Observation"]])
        out["log_action"] = np.stack([np.asarray(a, np.float64) for a in lg["Action"]])
        for col, key in (("Reward", "reward"), ("Cashflow", "cashflow"), ("Penalties", "penalties"),
                         ("Grid overloading", "overload"), ("SOC violation", "soc_viol")):
            out["log_" + key] = lg[col].to_numpy(np.float64)
        out["log_degradation"] = np.stack([np.broadcast_to(np.asarray(d, np.float64), (N,)) for d in lg["Degradation"]])
        out["log_deg_is_array"] = np.array([np.ndim(d) == 1 for d in lg["Degradation"]])
        out["log_charging_energy"] = np.stack([np.asarray(c, np.float64) for c in lg["Charging energy"]])
        out["log_soh"] = np.stack([np.asarray(h, np.float64) for h in lg["SOH"]])
    out["consts_json"] = np.array(json.dumps(consts))
    out["meta_json"] = np.array(json.dumps(dict(name=name, cfg={k: v for k, v in cfg.items() if k != "data_path"},
                                                starts=starts, window=[w0, w1], linear=linear, generated=generated,
                                                numpy=np.__version__, pandas=pd.__version__)))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: N={N} D={len(rec['obs'][0])} steps={len(rec['reward'])} "
          f"sum_reward={sum(rec['reward']):.6f} final_soh_min={rec['soh'][-1].min():.9f} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def uniform(rng, k, N, env):
    return rng.uniform(-1, 1, N)


def ones(rng, k, N, env):
    return np.ones(N)


def mixed(rng, k, N, env):
    """mostly charge, sometimes idle (exact zeros), sometimes discharge: exercises plateaus in the SOC history"""
    a = rng.uniform(-0.3, 1, N)
    a[rng.random(N) < 0.3] = 0.0
    return a


def mixed_pm(rng, k, N, env):
    """charging and discharging cars side by side in most steps, some idle"""
    a = rng.uniform(-1, 1, N)
    a[rng.random(N) < 0.15] = 0.0
    return a


ARBITRAGE = dict(spot_markup=0, spot_mul=1, feed_in_ded=0)
TARIFF = dict(tariff_name="fixed_feed_in.csv", spot_markup=10, spot_mul=1.5, feed_in_ded=0.25)

CASES = [
    # BASELINE cfg1: lmd, 1 EV, 15 min, one episode of random actions, linear degradation (wiring patch B-1)
    dict(name="cfg1_lmd_1ev_linear", cfg_over=dict(ARBITRAGE), starts=["2020-01-02 19:00"], n_steps_per_ep=96,
         action_fn=uniform, linear=True),
    dict(name="lmd_1ev_sei", cfg_over=dict(ARBITRAGE), starts=["2020-01-02 19:00"], n_steps_per_ep=96,
         action_fn=uniform),
    dict(name="lmd_5ev_24h", cfg_over=dict(ARBITRAGE), starts=["2020-03-10 06:15"], n_steps_per_ep=96,
         action_fn=uniform, n_evs=5, seed=1),
    dict(name="lmd_20ev_48h_mixed", cfg_over=dict(ARBITRAGE, episode_length=48), starts=["2020-05-04 12:00"],
         n_steps_per_ep=192, action_fn=mixed, n_evs=20, seed=2),
    dict(name="lmd_20ev_uncontrolled", cfg_over=dict(TARIFF), starts=["2020-02-03 15:30"], n_steps_per_ep=96,
         action_fn=ones, n_evs=20, seed=3),
    dict(name="ct_5ev_48h", cfg_over=dict(TARIFF, use_case="ct", schedule_name="1_ct.csv", building_name="load_ct.csv",
                                          episode_length=48),
         starts=["2020-06-08 05:00"], n_steps_per_ep=192, action_fn=uniform, n_evs=5, seed=4),
    dict(name="ut_5ev_1h_priceonly", cfg_over=dict(TARIFF, use_case="ut", schedule_name="1_ut.csv",
                                                   building_name="load_ut.csv", include_building=False,
                                                   include_pv=False, episode_length=48, freq="1H", minutes=60,
                                                   time_steps_per_hour=1),
         starts=["2020-09-07 03:00"], n_steps_per_ep=48, action_fn=uniform, n_evs=5, seed=5),
    dict(name="ut_5ev_15min_normalized", cfg_over=dict(TARIFF, use_case="ut", schedule_name="1_ut.csv",
                                                       building_name="load_ut.csv", normalize_in_env=True),
         starts=["2020-04-14 14:30"], n_steps_per_ep=96, action_fn=uniform, n_evs=5, seed=6),
    dict(name="lmd_5ev_buildingonly_noaux", cfg_over=dict(ARBITRAGE, include_pv=False, aux=False),
         starts=["2020-07-01 00:00"], n_steps_per_ep=96, action_fn=mixed, n_evs=5, seed=7),
    dict(name="lmd_5ev_pvonly", cfg_over=dict(ARBITRAGE, include_building=False), starts=["2020-07-02 09:45"],
         n_steps_per_ep=96, action_fn=uniform, n_evs=5, seed=8),
    # the same env object reused for three episodes: degradation state carried across resets (SURVEY B-3)
    dict(name="lmd_5ev_three_episodes", cfg_over=dict(ARBITRAGE),
         starts=["2020-03-02 08:00", "2020-03-03 08:00", "2020-03-02 20:15"], n_steps_per_ep=96,
         action_fn=uniform, n_evs=5, seed=9),
    # used battery: soh <= 0.9 flips target_soc to 0.9 during the first step (fleet_environment.py:613-614)
    dict(name="lmd_5ev_soh09_nodeg", cfg_over=dict(ARBITRAGE, init_soh=0.9, calculate_degradation=False),
         starts=["2020-03-10 06:15"], n_steps_per_ep=96, action_fn=uniform, n_evs=5, seed=10),
    # ---- BASELINE.json fleet shapes (cfg2 / cfg3 / cfg4): fleets from the product's generator, written to the reference's
    # CSV dialects (schedule.write_reference_csvs) and read back by the unmodified reference
    dict(name="cfg2_lmd_50ev_gen", cfg_over=dict(TARIFF, price_name=None, tariff_name=None), starts=["2020-03-10 06:15"],
         n_steps_per_ep=96, action_fn=uniform, n_evs=50, seed=11, generated=101),
    # SURVEY 8d cfg2 parity subset: sixteen 96-step episodes at the cfg2 fleet shape (every weekday and time of day, three weeks), on ONE env object
    # (degradation state carried from episode to episode like in an SB3 worker)
    dict(name="cfg2_lmd_50ev_gen_16ep", cfg_over=dict(TARIFF, price_name=None, tariff_name=None),
         starts=["2020-03-02 00:00", "2020-03-03 05:30", "2020-03-04 11:15", "2020-03-05 17:45", "2020-03-06 23:00",
                 "2020-03-08 08:15", "2020-03-09 14:00", "2020-03-10 20:30", "2020-03-12 02:45", "2020-03-13 09:00",
                 "2020-03-14 15:15", "2020-03-15 21:30", "2020-03-17 03:45", "2020-03-18 10:00", "2020-03-19 16:15",
                 "2020-03-20 22:30"],
         n_steps_per_ep=96, action_fn=mixed_pm, n_evs=50, seed=15, generated=101),
    dict(name="cfg3_ct_20ev_48h_gen", cfg_over=dict(TARIFF, use_case="ct", episode_length=48, price_name=None, tariff_name=None),
         starts=["2020-06-08 05:00"], n_steps_per_ep=192, action_fn=uniform, n_evs=20, seed=12, generated=102),
    dict(name="cfg4_ut_50ev_1h_priceonly_gen",
         cfg_over=dict(TARIFF, use_case="ut", include_building=False, include_pv=False, episode_length=48, freq="1H",
                       minutes=60, time_steps_per_hour=1, price_name=None, tariff_name=None),
         starts=["2020-09-07 03:00"], n_steps_per_ep=48, action_fn=uniform, n_evs=50, seed=13, generated=103),
    # ---- DataLogger: log_data=True, two episodes on one env object (Episode numbering, reset rows, the 14:45 Degradation
    # row, EvCharger's charge_log with its car-to-car carry-over, ev_charger.py:81-82,212)
    dict(name="lmd_5ev_log_two_episodes", cfg_over=dict(ARBITRAGE), starts=["2020-03-02 20:00", "2020-03-03 20:00"],
         n_steps_per_ep=96, action_fn=mixed_pm, n_evs=5, seed=14, log=True),
]


def main():
    if not compat.available():
        raise SystemExit("reference not available; golden files can only be generated in the build container")
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case["name"] not in only:
            continue
        run_case(**case)


if __name__ == "__main__":
    main()
