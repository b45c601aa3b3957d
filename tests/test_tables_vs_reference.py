"""Table builder (fleetrl_b200/tables.py) vs the live reference DataLoader, in the build container only.

Every column the step reads must be IDENTICAL (array_equal) to what the unmodified reference puts in its `db`
(data_processing.py) — including the monthly de-trended reward curves and the caretaker lunch fix.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.reference

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "refshim"))

CASES = {
    "lmd_1ev": dict(),
    "ct_1ev": dict(use_case="ct", schedule_name="1_ct.csv", building_name="load_ct.csv",
                   tariff_name="fixed_feed_in.csv", spot_markup=10, spot_mul=1.5, feed_in_ded=0.25),
    "ut_1ev_1h": dict(use_case="ut", schedule_name="1_ut.csv", building_name="load_ut.csv", freq="1H", minutes=60,
                      time_steps_per_hour=1, include_building=False, include_pv=False),
    "lmd_3ev_stacked": dict(_n_evs=3),
    "lkw_custom_2021": dict(use_case="custom", schedule_name="1_lkw.csv", price_name="spot_2021_new.csv",
                            tariff_name="spot_2021_new_tariff.csv", custom_ev_battery_size_in_kwh=600,
                            custom_ev_charger_power_in_kw=120, custom_grid_connection_in_kw=500,
                            init_battery_cap=600.0, obc_max_power=250.0, max_batt_cap_in_all_use_cases=600),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_builder_matches_reference_db(name):
    import compat
    from fleetrl_b200.tables import build_fleet

    over = dict(CASES[name])
    n_evs = over.pop("_n_evs", 1)
    cfg = compat.base_config(**over)
    if n_evs > 1:
        sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
        import gen_golden
        cfg["data_path"] = gen_golden.make_data_path(cfg["schedule_name"], n_evs, f"{n_evs}_tbl.csv")
        cfg["schedule_name"] = f"{n_evs}_tbl.csv"
    env = compat.make_reference_env(cfg)
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import gen_golden
    ref_consts = gen_golden.extract_consts(env, deg_mode=0)
    N = int(env.num_cars)
    T = len(env.db) // N
    ref_tables, _ = gen_golden.extract_tables(env, 0, T)

    built = build_fleet(cfg, auto_reset=False)
    mine = built.consts.to_dict()
    for k, v in ref_consts.items():
        if k in ("auto_reset", "start_lo", "start_hi", "seed", "carry_degradation_state"):
            continue
        assert mine[k] == v, f"const {k}: {mine[k]!r} != {v!r}"
    for k, v in ref_tables.items():
        if v is None:
            assert built.tables[k] is None
            continue
        np.testing.assert_array_equal(np.asarray(built.tables[k]), v, err_msg=f"table {k}")
    # static time picker index (static_time_picker.py): default start shifted to the db year
    obs, _ = env.reset()
    t0 = int(np.searchsorted(built.dates, np.datetime64(env.episode.time)))
    assert built.start_ranges["static"][0] == t0


@pytest.mark.parametrize("name", ["lmd_1ev", "ut_1ev_1h", "ct_1ev"])
@pytest.mark.parametrize("end_cutoff", [60, 10])
def test_time_picker_candidate_ranges_match_reference(name, end_cutoff, monkeypatch):
    """tables.start_index_ranges vs the populations the UNMODIFIED pickers draw from: random_time_picker.py:25-28 and
    eval_time_picker.py:33-36 (random.choice is intercepted, the candidate DatetimeIndex it receives is compared element
    by element with dates[lo .. hi])."""
    import random

    import compat
    from fleetrl_b200.tables import build_fleet

    cfg = compat.base_config(**{k: v for k, v in CASES[name].items() if not k.startswith("_")})
    cfg["end_cutoff"] = end_cutoff
    env = compat.make_reference_env(cfg)
    from fleetrl.utils.time_picker.eval_time_picker import EvalTimePicker
    from fleetrl.utils.time_picker.random_time_picker import RandomTimePicker

    built = build_fleet(cfg, auto_reset=False)
    seen = {}

    def fake_choice(seq):
        seen["pop"] = seq
        return seq[0]

    monkeypatch.setattr(random, "choice", fake_choice)
    for key, picker in (("random", RandomTimePicker()), ("eval", EvalTimePicker(env.time_conf.episode_length))):
        picker.choose_time(env.db, env.time_conf.freq, env.time_conf.end_cutoff)
        pop = np.asarray(seen["pop"].values, dtype="datetime64[ns]")
        lo, hi = built.start_ranges[key]
        assert hi - lo + 1 == len(pop), f"{key}: {hi - lo + 1} candidates, reference has {len(pop)}"
        np.testing.assert_array_equal(built.dates[lo:hi + 1].astype("datetime64[ns]"), pop, err_msg=key)
    # and what reaches the device RNG: the configured picker's range
    for key in ("random", "eval"):
        b2 = build_fleet(dict(cfg, time_picker=key), auto_reset=True)
        assert (b2.consts.start_lo, b2.consts.start_hi) == tuple(built.start_ranges[key])
