"""Seeded random configurations through the default kernel selection (GPU vs oracle, via the C ABI): vehicle counts on
both sides of every kernel-selection threshold, odd / even N, all observer variants, normalisation, aux on / off, the
three degradation settings, auto-reset on / off, 15-min and 1-h grids.  Same tolerances as test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleFleet
from synth_tables import make_consts, make_tables

pytestmark = pytest.mark.gpu


def _draw(seed):
    r = np.random.default_rng(1000 + seed)
    n_evs = int(r.choice([1, 2, 7, 8, 9, 16, 31, 33, 64, 100, 129, 257]))
    use_case = str(r.choice(["lmd", "ut", "ct"]))
    sph = int(r.choice([4, 4, 1]))
    over = dict(normalize=int(r.integers(0, 2)), aux=int(r.integers(0, 2)), auto_reset=int(r.random() < 0.8),
                carry_degradation_state=int(r.integers(0, 2)))
    deg = int(r.integers(0, 3))
    if deg == 0:
        over["calc_degradation"] = 0
    elif deg == 1:
        over["deg_mode"] = 1
    obs_kind = int(r.integers(0, 4))          # full / price-only / price+load / price+pv
    if obs_kind == 1:
        over.update(include_building=0, include_pv=0)
    elif obs_kind == 2:
        over.update(include_pv=0)
    elif obs_kind == 3:
        over.update(include_building=0)
        over["normalize"] = 0                  # PV-only + normalisation raises in the reference
    if r.random() < 0.3:
        over["init_soh"] = 0.9
    E = int(r.choice([1, 5, 33, 64]))
    if n_evs >= 100:
        E = min(E, 5)
    return dict(n_evs=n_evs, use_case=use_case, sph=sph, over=over, E=E, two_trips=(use_case == "ct"),
                episode_hours=int(r.choice([24, 48])), steps=70 if sph == 1 else 130)


@pytest.mark.parametrize("seed", range(32))
def test_random_configuration(seed, monkeypatch):
    from fleetrl_b200._lib import FleetStepHandle
    monkeypatch.delenv("FLEETSTEP_KERNEL", raising=False)
    # odd seeds run with a tiny history ring / inline rainflow stack (ring flushes, extension slots)
    if seed % 2:
        monkeypatch.setenv("FLEETSTEP_RF_RING", "4" if seed % 4 == 1 else "8")
        monkeypatch.setenv("FLEETSTEP_RF_STACK", "2" if seed % 4 == 1 else "5")
        monkeypatch.setenv("FLEETSTEP_RF_EXT_SLOTS", "40000")      # one extension slot per vehicle is possible
    else:
        monkeypatch.delenv("FLEETSTEP_RF_RING", raising=False)
        monkeypatch.delenv("FLEETSTEP_RF_STACK", raising=False)
        monkeypatch.delenv("FLEETSTEP_RF_EXT_SLOTS", raising=False)
    cf = _draw(seed)
    N, E, over = cf["n_evs"], cf["E"], cf["over"]
    cap0 = dict(lmd=60.0, ut=50.0, ct=16.7)[cf["use_case"]]
    tables, T = make_tables(seed=seed, n_evs=N, days=30 if cf["sph"] == 1 else 10, sph=cf["sph"], two_trips=cf["two_trips"],
                            cap=cap0)
    if not over.get("include_building", 1):
        tables = dict(tables, load=None)
    if not over.get("include_pv", 1):
        tables = dict(tables, pv=None)
    consts = make_consts(tables, T, N, sph=cf["sph"], episode_hours=cf["episode_hours"], use_case=cf["use_case"], **over)
    orc = OracleFleet(consts, tables, E, env_id_offset=7)
    gpu = FleetStepHandle(consts, tables, E, device=0, env_id_offset=7)
    dev, D = gpu.device, gpu.D
    assert D == orc.D
    obs = torch.zeros((E, D), dtype=torch.float32, device=dev)
    term = torch.zeros((E, D), dtype=torch.float32, device=dev)
    rew = torch.zeros(E, dtype=torch.float32, device=dev)
    done = torch.zeros(E, dtype=torch.uint8, device=dev)
    o_obs = orc.reset()
    gpu.reset(obs=obs)
    np.testing.assert_array_equal(obs.cpu().numpy(), o_obs)
    rng = np.random.default_rng(seed)
    soc_exact = (not consts.calc_degradation) or consts.deg_mode == 1
    for s in range(cf["steps"]):
        a = rng.uniform(-1, 1, (E, N)).astype(np.float32)
        if s % 6 == 0:
            a[rng.random((E, N)) < 0.3] = 0.0
        o_obs, o_rew, o_cash, o_done, o_term = orc.step(a, want_terminal=True)
        gpu.step(torch.from_numpy(a).to(dev), obs, rew, done, term)
        msg = f"{cf} step {s}"
        np.testing.assert_array_equal(done.cpu().numpy(), o_done, err_msg=msg)
        for k in ("time_idx", "hours_left", "target_soc", "rf_len", "n_cycles"):
            np.testing.assert_array_equal(gpu.get(k).cpu().numpy(), orc.get(k), err_msg=f"{k} {msg}")
        g_soc, o_soc = gpu.get("soc").cpu().numpy(), orc.get("soc")
        if soc_exact:
            np.testing.assert_array_equal(g_soc, o_soc, err_msg=msg)
        else:
            np.testing.assert_allclose(g_soc, o_soc, rtol=0, atol=1e-12, err_msg=msg)
        np.testing.assert_array_equal(obs.cpu().numpy(), o_obs, err_msg=msg)
        np.testing.assert_allclose(gpu.get("reward64").cpu().numpy(), o_rew, rtol=1e-11, atol=1e-10, err_msg=msg)
        np.testing.assert_allclose(gpu.get("cashflow").cpu().numpy(), o_cash, rtol=1e-12, atol=1e-13, err_msg=msg)
        np.testing.assert_allclose(gpu.get("soh").cpu().numpy(), orc.get("soh"), rtol=0, atol=1e-13, err_msg=msg)
        if consts.auto_reset and o_done.any():
            idx = np.nonzero(o_done)[0]
            np.testing.assert_array_equal(term.cpu().numpy()[idx], o_term[idx], err_msg=msg)
    g_stats, o_stats = gpu.stats(), orc.stats()
    for k in o_stats:
        np.testing.assert_allclose(g_stats[k], o_stats[k], rtol=1e-9, atol=1e-9, err_msg=f"stat {k} {cf}")
    assert gpu.check_errors() == 0
    gpu.close()
