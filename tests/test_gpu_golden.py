"""GPU parity against the reference's golden trajectories, through the C ABI (fleetrl_b200/libfleetstep.so).

Same fixtures and the same tolerances as tests/test_oracle_golden.py:
  bit-exact  : done, hours_left, target_soc, rainflow_length (cycle counts), soc, soc_deg, observation (float32)
  rel 1e-11 / abs 1e-10 : reward (per-vehicle terms are added per vehicle first and then summed over the
               fleet in car order — a deterministic re-association of the reference's sum — plus device exp vs numpy exp);
  rel 1e-12  : cashflow (revenue factor regrouped)
  abs 1e-13  : SOH, l ; rel 1e-12 : fd_cyc   (device pow/exp vs numpy, summation order of the stress terms)
"""
import numpy as np
import pytest
import torch

from golden_util import Golden, golden_names

pytestmark = pytest.mark.gpu


def run_gpu(g: Golden, **const_over):
    from fleetrl_b200._lib import FleetStepHandle

    consts = g.consts(**const_over)
    h = FleetStepHandle(consts, g.tables, num_envs=1, device=0)
    dev = h.device
    keys = ["soc", "hours_left", "soc_deg", "soh", "target_soc", "rf_len", "fd_cyc", "life"]
    out = {k: [] for k in ["obs", "reward", "cashflow", "done"] + keys}
    obs = torch.zeros((1, h.D), dtype=torch.float32, device=dev)
    rew = torch.zeros(1, dtype=torch.float32, device=dev)
    done = torch.zeros(1, dtype=torch.uint8, device=dev)

    def snap():
        out["obs"].append(obs[0].cpu().numpy().copy())
        for k in keys:
            out[k].append(h.get(k)[0].cpu().numpy().copy())

    step = 0
    for ep, t0 in enumerate(g.start_idx):
        h.reset(start_idx=torch.tensor([int(t0)], dtype=torch.int32, device=dev), obs=obs)
        snap()
        for k in range(g.n_steps_per_ep):
            a = torch.from_numpy(g.actions[step][None, :].copy()).to(dev)
            h.step(a, obs, rew, done)
            step += 1
            out["reward"].append(h.get("reward64")[0].item())
            out["cashflow"].append(h.get("cashflow")[0].item())
            out["done"].append(bool(done[0].item()))
            snap()
    assert h.check_errors() == 0
    h.close()
    return {k: np.array(v) for k, v in out.items()}


@pytest.mark.parametrize("name", golden_names())
def test_gpu_matches_reference(name):
    g = Golden(name)
    o = run_gpu(g)
    r = g.traj
    assert o["obs"].shape == r["obs"].shape
    np.testing.assert_array_equal(o["done"], r["done"])
    np.testing.assert_array_equal(o["hours_left"], r["hours_left"].astype(np.float32))
    np.testing.assert_array_equal(o["target_soc"], r["target_soc"])
    np.testing.assert_array_equal(o["soc"], r["soc"])
    np.testing.assert_array_equal(o["soc_deg"], r["soc_deg"])
    np.testing.assert_allclose(o["reward"], r["reward"], rtol=1e-11, atol=1e-10)
    np.testing.assert_allclose(o["cashflow"], r["cashflow"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(o["soh"], r["soh"], rtol=0, atol=1e-13)
    if "rf_len" in r:
        np.testing.assert_array_equal(o["rf_len"], r["rf_len"].astype(np.int32))
        np.testing.assert_allclose(o["fd_cyc"], r["fd_cyc"], rtol=1e-12, atol=1e-18)
        np.testing.assert_allclose(o["life"], r["life"], rtol=0, atol=1e-13)
    np.testing.assert_array_equal(o["obs"], r["obs"])


@pytest.mark.parametrize("name", [n for n in golden_names() if "_log_" in n])
def test_gpu_log_matches_reference_datalogger(name):
    """The device-side log ring (fleet_enable_log) against the reference's own DataLogger frame (log_data=True): every
    column of data_logger.py:55-68, including the Episode counter, the reset rows, the 14:45 Degradation row and
    EvCharger's charge_log with its car-to-car carry-over (ev_charger.py:81-82,212)."""
    from fleetrl_b200._lib import FleetStepHandle

    g = Golden(name)
    ref = g.log
    h = FleetStepHandle(g.consts(), g.tables, num_envs=1, device=0)
    dev = h.device
    obs = torch.zeros((1, h.D), dtype=torch.float32, device=dev)
    rew = torch.zeros(1, dtype=torch.float32, device=dev)
    done = torch.zeros(1, dtype=torch.uint8, device=dev)
    h.enable_log([0], len(ref["reward"]) + 8)
    step = 0
    for ep, t0 in enumerate(g.start_idx):
        h.reset(start_idx=torch.tensor([int(t0)], dtype=torch.int32, device=dev), obs=obs)
        for k in range(g.n_steps_per_ep):
            h.step(torch.from_numpy(g.actions[step][None, :].copy()).to(dev), obs, rew, done)
            step += 1
    rec = h.read_log()[0]
    assert h.check_errors() == 0
    h.close()
    n = len(ref["reward"])
    assert rec["rows_total"] == n == len(rec["kind"])
    L = g.n_steps_per_ep
    np.testing.assert_array_equal(np.arange(n) // L + 1, ref["episode"])                 # data_logger.py:51
    np.testing.assert_array_equal(rec["time_idx"], ref["time_idx"])
    np.testing.assert_array_equal(rec["obs"], ref["obs"])
    np.testing.assert_array_equal(rec["action"], ref["action"])
    np.testing.assert_allclose(rec["reward"], ref["reward"], rtol=1e-11, atol=1e-10)
    np.testing.assert_allclose(rec["cashflow"], ref["cashflow"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(rec["penalties"], ref["penalties"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(rec["overload"], ref["overload"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(rec["soc_viol"], ref["soc_viol"], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(rec["kind"] == 2, ref["deg_is_array"])
    np.testing.assert_allclose(rec["degradation"], ref["degradation"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(rec["charging_energy"], ref["charging_energy"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(rec["soh"], ref["soh"], rtol=0, atol=1e-13)
