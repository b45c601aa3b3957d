"""Pin the CPU oracle (oracle/fleet_oracle.c) against trajectories of the UNMODIFIED reference FleetEnv.

The fixtures under tests/golden/ were produced in the build container by oracle/gen_golden.py.

Tolerances (stated):
  - bit-exact: done, time index, hours_left, target_soc, rainflow_length (cycle counts), soc, soc_deg
    (per-car float64 in the reference's operation order, IEEE ops only)
  - reward / cashflow: rel 1e-12 (libm exp vs numpy exp in the two sigmoid penalties; everything else is the
    same sequence of IEEE operations)
  - SOH / fd_cyc / l: abs 1e-13 (libm pow vs numpy power, pandas' summation order)
  - obs (float32): bit-exact, except elements fed by exp/pow/sum paths (none are) -> asserted exact
"""
import numpy as np
import pytest

from golden_util import Golden, golden_names
from oracle.oracle import OracleFleet


def run_oracle(g: Golden, **const_over):
    consts = g.consts(**const_over)
    orc = OracleFleet(consts, g.tables, num_envs=1)
    out = {k: [] for k in ["obs", "reward", "cashflow", "done", "soc", "hours_left", "soc_deg", "soh", "target_soc",
                           "rf_len", "fd_cyc", "life"]}

    def snap(obs):
        out["obs"].append(obs[0].copy())
        for k in ["soc", "hours_left", "soc_deg", "soh", "target_soc", "rf_len", "fd_cyc", "life"]:
            out[k].append(orc.get(k)[0].copy())

    step = 0
    for ep, t0 in enumerate(g.start_idx):
        snap(orc.reset(start_idx=[t0]))
        for k in range(g.n_steps_per_ep):
            obs, rew, cash, done = orc.step(g.actions[step][None, :])
            step += 1
            out["reward"].append(rew[0]); out["cashflow"].append(cash[0]); out["done"].append(bool(done[0]))
            snap(obs)
    assert orc.err_flags() == 0
    return {k: np.array(v) for k, v in out.items()}


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    g = Golden(name)
    o = run_oracle(g)
    r = g.traj
    assert o["obs"].shape == r["obs"].shape
    np.testing.assert_array_equal(o["done"], r["done"])
    np.testing.assert_array_equal(o["hours_left"], r["hours_left"].astype(np.float32))
    np.testing.assert_array_equal(o["target_soc"], r["target_soc"])
    np.testing.assert_array_equal(o["soc"], r["soc"])
    np.testing.assert_array_equal(o["soc_deg"], r["soc_deg"])
    np.testing.assert_allclose(o["reward"], r["reward"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(o["cashflow"], r["cashflow"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(o["soh"], r["soh"], rtol=0, atol=1e-13)
    if "rf_len" in r:
        np.testing.assert_array_equal(o["rf_len"], r["rf_len"].astype(np.int32))
        np.testing.assert_allclose(o["fd_cyc"], r["fd_cyc"], rtol=1e-12, atol=1e-18)
        np.testing.assert_allclose(o["life"], r["life"], rtol=0, atol=1e-13)
    np.testing.assert_array_equal(o["obs"], r["obs"])
