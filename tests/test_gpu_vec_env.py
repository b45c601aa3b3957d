"""The SB3-shaped FleetVecEnv and the gym-shaped FleetEnv on the GPU, checked against the oracle through the
public API (synthetic fleet from the product's own generator + table builder)."""
import numpy as np
import pandas as pd
import pytest
import torch

from fleetrl_b200.config import default_config
from fleetrl_b200.schedule import generate_schedule, synthetic_series
from fleetrl_b200.tables import FleetInputs, build_fleet
from oracle.oracle import OracleFleet

pytestmark = pytest.mark.gpu


def _inputs(use_case="lmd", n=6):
    sched = generate_schedule(use_case, n, start="2020-01-01 00:00", end="2020-03-31 23:59", seed=11)
    price, tariff, load, pv = synthetic_series(start="2020-01-01 00:00", end="2020-03-31 23:59")
    return FleetInputs(sched, price, tariff, load, pv)


@pytest.mark.parametrize("output", ["torch", "numpy"])
def test_vec_env_matches_oracle(output):
    from fleetrl_b200 import FleetVecEnv
    cfg = default_config("lmd", time_picker="random", end_cutoff=10)
    inputs = _inputs()
    E = 48
    env = FleetVecEnv(cfg, E, inputs=inputs, output=output, env_id_offset=77, seed=3)
    orc = OracleFleet(env.built.consts, env.built.tables, E, env_id_offset=77)
    assert env.observation_space.shape == (env.obs_dim,) and env.action_space.shape == (6,)
    obs = env.reset()
    o_obs = orc.reset()
    np.testing.assert_array_equal(obs.cpu().numpy() if output == "torch" else obs, o_obs)
    rng = np.random.default_rng(0)
    seen_done = 0
    for s in range(200):
        a = rng.uniform(-1, 1, (E, 6)).astype(np.float32)
        obs, rew, done, infos = env.step(a if output == "numpy" else torch.from_numpy(a).to(env.device))
        o_obs, o_rew, _, o_done, o_term = orc.step(a, want_terminal=True)
        g_obs = obs.cpu().numpy() if output == "torch" else obs
        g_done = done.cpu().numpy() if output == "torch" else done
        g_rew = rew.cpu().numpy() if output == "torch" else rew
        np.testing.assert_array_equal(g_done, o_done.astype(bool))
        np.testing.assert_array_equal(g_obs, o_obs)
        np.testing.assert_allclose(g_rew, o_rew.astype(np.float32), rtol=1e-6, atol=1e-6)
        assert len(infos) == E
        for i in np.nonzero(o_done)[0]:
            info = infos[int(i)]
            t = info["terminal_observation"]
            np.testing.assert_array_equal(t.cpu().numpy() if output == "torch" else t, o_term[i])
            assert info["TimeLimit.truncated"] is False and info["episode"]["l"] == 96
            np.testing.assert_allclose(info["episode"]["r"], orc.get("last_ep_return")[i], rtol=1e-11)
            seen_done += 1
        if not o_done.any():
            assert infos[0] == {}
    assert seen_done >= E
    # env_method surface
    assert env.env_method("is_done", indices=[0]) == [bool(o_done[0])]
    t = env.env_method("get_time", indices=0)[0]
    assert t == env.built.dates[orc.get("time_idx")[0]]
    df = env.env_method("get_dist_factor", indices=[1])[0]
    assert df.shape == (6,)
    assert env.env_is_wrapped(None) == [False] * E
    st, ost = env.stats(), orc.stats()
    for k in ost:
        np.testing.assert_allclose(st[k], ost[k], rtol=1e-9, atol=1e-9, err_msg=k)
    env.close()


def FleetVecEnv_LOG_COLUMNS():
    from fleetrl_b200 import FleetVecEnv
    return FleetVecEnv.LOG_COLUMNS


def test_fleet_env_log_data_flag():
    """env_config["log_data"]=True (fleet_environment.py:129): FleetEnv.get_log() is the DataLogger frame — one reset row and
    one row per step that does not end the episode (:420-432, :679-690)."""
    from fleetrl_b200 import FleetEnv
    cfg = default_config("lmd", time_picker="static", episode_length=24, log_data=True)
    env = FleetEnv(cfg, inputs=_inputs("lmd", 3))
    env.reset()
    rng = np.random.default_rng(3)
    rewards = []
    for s in range(96):
        _, r, done, _, _ = env.step(rng.uniform(-1, 1, 3).astype(np.float32))
        rewards.append(r)
    assert done
    log = env.get_log()
    assert list(log.columns) == list(FleetVecEnv_LOG_COLUMNS()) and len(log) == 96      # reset row + 95 non-finishing steps
    np.testing.assert_allclose(log["Reward"].to_numpy(dtype=float)[1:], rewards[:-1], rtol=1e-12, atol=1e-12)
    assert (log["Episode"] == 1).all() and log["Time"].iloc[1] - log["Time"].iloc[0] == pd.Timedelta(minutes=15)
    env.close()


def test_fleet_env_gym_api():
    from fleetrl_b200 import FleetEnv
    cfg = default_config("ct", time_picker="static", episode_length=48)
    inputs = _inputs("ct", 4)
    env = FleetEnv(cfg, inputs=inputs)
    built = build_fleet(cfg, inputs, auto_reset=False)
    orc = OracleFleet(built.consts, built.tables, 1)
    obs, info = env.reset()
    assert info == {} and obs.dtype == np.float32 and obs.shape == env.observation_space.shape
    t0 = built.start_ranges["static"][0]
    np.testing.assert_array_equal(obs, orc.reset(start_idx=[t0])[0])
    assert env.get_time() == built.dates[t0] == env.get_start_time()
    rng = np.random.default_rng(1)
    for s in range(192):
        a = rng.uniform(-1, 1, 4).astype(np.float32)
        obs, r, done, trunc, info = env.step(a)
        o_obs, o_r, _, o_d = orc.step(a[None, :])
        np.testing.assert_array_equal(obs, o_obs[0])
        assert isinstance(r, float) and trunc is False and info == {}
        np.testing.assert_allclose(r, o_r[0], rtol=1e-11, atol=1e-10)
        assert done == bool(o_d[0])
    assert done and env.is_done()
    assert list(env.get_log().columns) != list(FleetVecEnv_LOG_COLUMNS())          # log_data=False: reduced statistics
    with pytest.raises(TypeError):
        env.step(np.array([np.nan, 0, 0, 0], dtype=np.float32))
    env.close()


def _carry_over(energy, actions):
    """EvCharger's charge_log entry is charging_energy + discharging_energy, and those two locals survive from car to car
    (ev_charger.py:81-82,212): a car's entry includes the last opposite-sign car's energy (pinned against the reference by
    the log golden, tests/test_gpu_golden.py)."""
    out, lc, ld = np.zeros(len(energy)), 0.0, 0.0
    for n in range(len(energy)):
        if actions[n] >= 0:
            lc = energy[n]
        else:
            ld = energy[n]
        out[n] = lc + ld
    return out


@pytest.mark.parametrize("output", ["torch", "numpy"])
def test_evaluation_log_columns_and_values(output):
    """enable_log(): the reference's DataLogger rows (reset row, one row per non-terminal step) for chosen envs, written
    by the device-side log ring (fleet_enable_log) in both output modes; every column is checked against the oracle."""
    from fleetrl_b200 import FleetVecEnv
    cfg = default_config("lmd", time_picker="random", end_cutoff=10)
    E, N = 8, 6
    env = FleetVecEnv(cfg, E, inputs=_inputs(), output=output, env_id_offset=5, seed=3)
    orc = OracleFleet(env.built.consts, env.built.tables, E, env_id_offset=5)
    env.enable_log(indices=[0, 3])
    env.reset(); o_obs = orc.reset()
    pm = float(env.built.consts.price_multiplier)
    rng = np.random.default_rng(2)
    want = {i: [dict(kind="reset", obs=o_obs[i].copy(), soh=orc.get("soh")[i].copy())] for i in (0, 3)}
    for s in range(130):
        a = rng.uniform(-1, 1, (E, N)).astype(np.float32)
        a[rng.random((E, N)) < 0.1] = 0.0
        env.step(a if output == "numpy" else torch.from_numpy(a).to(env.device))
        o_obs, o_rew, o_cash, o_done, _ = orc.step(a, want_terminal=True)
        for i in (0, 3):
            if o_done[i]:      # the finishing step is not logged; the auto-reset's row is (fleet_environment.py:420-432,679)
                want[i].append(dict(kind="reset", obs=o_obs[i].copy(), soh=orc.get("soh")[i].copy()))
            else:
                want[i].append(dict(kind="step", obs=o_obs[i].copy(), soh=orc.get("soh")[i].copy(), action=a[i].copy(),
                                    reward=o_rew[i], cash=o_cash[i], overload=orc.get("overload")[i],
                                    soc_viol=orc.get("soc_viol")[i], time_idx=orc.get("time_idx")[i],
                                    charge_log=_carry_over(orc.get("charge_log")[i], a[i]), last_deg=orc.get("last_deg")[i].copy()))
    for i in (0, 3):
        log = env.env_method("get_log", indices=[i])[0]
        assert list(log.columns) == list(env.LOG_COLUMNS)
        assert len(log) == len(want[i]) == 130 + 1          # every step gives one row: its own, or the reset row after a done
        assert log["Episode"].iloc[0] == 1 and log["Episode"].iloc[-1] == 2
        n_deg = 0
        for r, w in enumerate(want[i]):
            row = log.iloc[r]
            np.testing.assert_array_equal(row["Observation"], w["obs"], err_msg=f"row {r}")
            np.testing.assert_allclose(row["SOH"], w["soh"], rtol=0, atol=1e-13)
            if w["kind"] == "reset":
                assert row["Reward"] == 0.0 and row["Cashflow"] == 0.0 and row["Penalties"] == 0.0
                assert not np.any(row["Action"]) and not np.any(row["Charging energy"]) and np.ndim(row["Degradation"]) == 0
                continue
            assert row["Time"] == pd.Timestamp(env.built.dates[int(w["time_idx"])])
            np.testing.assert_array_equal(row["Action"], w["action"].astype(np.float64))
            np.testing.assert_allclose(row["Reward"], w["reward"], rtol=1e-11, atol=1e-10)
            np.testing.assert_allclose(row["Cashflow"], w["cash"], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(row["Penalties"], w["reward"] - w["cash"] * pm, rtol=1e-10, atol=1e-9)   # :659
            np.testing.assert_allclose(row["Grid overloading"], w["overload"], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(row["SOC violation"], w["soc_viol"], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(row["Charging energy"], w["charge_log"], rtol=1e-13, atol=1e-13)
            if np.ndim(row["Degradation"]) == 1:          # the 14:45 row carries the per-vehicle degradation (:665-676)
                np.testing.assert_allclose(row["Degradation"], w["last_deg"], rtol=0, atol=1e-13)
                ts = row["Time"]
                assert ts.hour == 14 and ts.minute == 45
                n_deg += 1
            else:
                assert row["Degradation"] == 0.0
        assert n_deg >= 1
    env.close()


def test_numpy_outputs_are_fresh_arrays():
    """SB3 keeps obs_t across env.step (rollout_buffer.add(self._last_obs, ...)): the arrays returned by step() must not be
    overwritten by the next step (DummyVecEnv / SubprocVecEnv return fresh arrays)."""
    from fleetrl_b200 import FleetVecEnv
    cfg = default_config("lmd", time_picker="random", end_cutoff=10)
    E, N = 16, 6
    env = FleetVecEnv(cfg, E, inputs=_inputs(), output="numpy", seed=3)
    env.reset()
    rng = np.random.default_rng(4)
    obs1, rew1, done1, _ = env.step(rng.uniform(-1, 1, (E, N)).astype(np.float32))
    keep = obs1.copy(), rew1.copy(), done1.copy()
    obs2, rew2, done2, _ = env.step(rng.uniform(-1, 1, (E, N)).astype(np.float32))
    np.testing.assert_array_equal(obs1, keep[0]); np.testing.assert_array_equal(rew1, keep[1])
    np.testing.assert_array_equal(done1, keep[2])
    assert not np.array_equal(obs1, obs2)
    env.close()


def test_state_dict_round_trip_resumes_bit_for_bit():
    """state_dict() / load_state_dict(): a second env object restored from the checkpoint continues the trajectories of the
    first one exactly (history ring, rainflow stacks, degradation members, counters and statistics included)."""
    from fleetrl_b200 import FleetVecEnv
    cfg = default_config("lmd", time_picker="random", end_cutoff=10)
    E, N = 32, 6
    inputs = _inputs()
    a_env = FleetVecEnv(cfg, E, inputs=inputs, output="torch", env_id_offset=9, seed=3)
    a_env.reset()
    rng = np.random.default_rng(6)
    acts = [torch.from_numpy(rng.uniform(-1, 1, (E, N)).astype(np.float32)).to(a_env.device) for _ in range(260)]
    for s in range(130):                       # past one episode end and one daily evaluation
        a_env.step(acts[s])
    sd = a_env.state_dict()
    b_env = FleetVecEnv(cfg, E, inputs=inputs, output="torch", env_id_offset=9, seed=3)
    obs_b = b_env.load_state_dict(sd)
    np.testing.assert_array_equal(obs_b.cpu().numpy(), a_env._obs.cpu().numpy())
    for s in range(130, 260):
        oa, ra, da, _ = a_env.step(acts[s])
        ob, rb, db, _ = b_env.step(acts[s])
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), f"step {s}"
    for k in ("soc", "soh", "hours_left", "rf_len", "fd_cyc", "life", "time_idx", "ep_return"):
        assert torch.equal(a_env.handle.get(k), b_env.handle.get(k)), k
    sa, sb = a_env.stats(), b_env.stats()
    for k in sa:
        assert sa[k] == sb[k], k
    with pytest.raises(Exception):
        FleetVecEnv(cfg, E + 1, inputs=inputs, output="torch", seed=3).load_state_dict(sd)
    a_env.close(); b_env.close()


def test_step_host_pageable_and_pinned_agree():
    """fleet_step_host with pageable NumPy buffers (staging copies) and with page-locked ones (used in place by the
    kernels) gives the same outputs as the device-buffer call."""
    from fleetrl_b200._lib import FleetStepHandle
    cfg = default_config("lmd", time_picker="random", end_cutoff=10)
    built = build_fleet(cfg, _inputs(), auto_reset=True)
    E, N = 40, 6
    hs = [FleetStepHandle(built.consts, built.tables, E, device=0, env_id_offset=3) for _ in range(3)]
    D = hs[0].D
    dev = hs[0].device
    obs_d = torch.zeros((E, D), dtype=torch.float32, device=dev)
    rew_d = torch.zeros(E, dtype=torch.float32, device=dev); done_d = torch.zeros(E, dtype=torch.uint8, device=dev)
    for h in hs:
        h.reset()
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    o_pin, r_pin, d_pin, a_pin = pin((E, D), torch.float32), pin(E, torch.float32), pin(E, torch.uint8), pin((E, N), torch.float32)
    o_pg, r_pg, d_pg = np.empty((E, D), np.float32), np.empty(E, np.float32), np.empty(E, np.uint8)
    rng = np.random.default_rng(5)
    for s in range(110):
        a = rng.uniform(-1, 1, (E, N)).astype(np.float32)
        hs[0].step(torch.from_numpy(a).to(dev), obs_d, rew_d, done_d)
        hs[1].step_host(a, o_pg, r_pg, d_pg)                    # pageable
        a_pin[...] = a
        hs[2].step_host(a_pin, o_pin, r_pin, d_pin)             # page-locked, in place
        for o, r, d in ((o_pg, r_pg, d_pg), (o_pin, r_pin, d_pin)):
            np.testing.assert_array_equal(o, obs_d.cpu().numpy(), err_msg=f"step {s}")
            np.testing.assert_array_equal(r, rew_d.cpu().numpy())
            np.testing.assert_array_equal(d, done_d.cpu().numpy())
    for h in hs:
        h.close()
