"""CPU-only tests of the host side: config mirror, spaces, lazy infos, sharding, the schedule generator + table
builder invariants, and that the C-ABI library loads and exports every symbol include/fleetstep.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from fleetrl_b200 import config as cfgmod
from fleetrl_b200._abi import FleetConsts
from fleetrl_b200.dist import shard_range
from fleetrl_b200.schedule import generate_schedule, synthetic_series
from fleetrl_b200.spaces import action_box, observation_box
from fleetrl_b200.tables import FleetInputs, build_fleet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    """Every function declared in include/fleetstep.h is exported by the built library (no compute calls)."""
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "fleetstep.h")).read()
    names = set(re.findall(r"\b(fleet_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 19
    lib = ctypes.CDLL(os.path.join(ROOT, "fleetrl_b200", "libfleetstep.so"))
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} is declared in fleetstep.h but not exported"
    lib.fleet_abi_version.restype = ctypes.c_int
    from fleetrl_b200._abi import ABI_VERSION
    assert lib.fleet_abi_version() == ABI_VERSION == 2


def test_consts_struct_matches_header_size():
    # 22 int32 + uint64 + 31 doubles, laid out like the C struct
    assert ctypes.sizeof(FleetConsts) == 88 + 8 + 8 * 31


def test_config_resolution_order():
    cfg = cfgmod.default_config("ct", spot_markup=0, spot_mul=1, feed_in_ded=0, target_soc=0.8,
                                max_batt_cap_in_all_use_cases=60, ignore_invalid_penalty=True)
    rc = cfgmod.resolve(cfg)
    assert rc.ev.init_battery_cap == 16.7 and rc.ev.fixed_markup == 0 and rc.ev.variable_multiplier == 1
    assert rc.ev.feed_in_deduction == 0 and rc.ev.target_soc == 0.8
    assert rc.score.price_multiplier == 3.33 * (60 / 16.7)          # fleet_environment.py:194
    assert rc.score.penalty_invalid_action == 0
    co = cfgmod.company_for("ct", cfg, max_load=100.0, num_cars=20)
    assert co.evse_max_power == 4.6 and co.grid_connection == max(110.00000000000001, 100 + 0.5 * 20 * 4.6)
    assert cfgmod.company_for("ut", cfg, 50.0, 5).grid_connection == 1000


def test_config_errors():
    cfg = cfgmod.default_config("lmd")
    bad = dict(cfg); del bad["target_soc"]
    with pytest.raises(KeyError):
        cfgmod.resolve(bad)
    with pytest.raises(NotImplementedError):
        cfgmod.resolve(dict(cfg, include_price=False))
    with pytest.raises(NotImplementedError):
        cfgmod.resolve(dict(cfg, real_time=True))
    with pytest.raises(TypeError):
        cfgmod.resolve(dict(cfg, use_case="bus"))
    with pytest.raises(AssertionError):
        cfgmod.read_config(3.14)


def test_spaces():
    ob = observation_box(388, normalized=False)
    assert ob.shape == (388,) and ob.dtype == np.float32 and np.isinf(ob.low).all()
    ob = observation_box(45, normalized=True)
    assert ob.low.min() == 0 and ob.high.max() == 1
    ab = action_box(50)
    assert ab.shape == (50,) and ab.low.min() == -1 and ab.high.max() == 1


def test_shard_range_covers_everything():
    for total, world in [(65536, 8), (1048576, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 4, 4)


def test_lazy_infos():
    from fleetrl_b200.vec_env import LazyInfos
    term = np.arange(6, dtype=np.float32).reshape(2, 3)
    infos = LazyInfos(5, [1, 4], term, np.array([-3.0, 2.5]), 96)
    assert len(infos) == 5 and infos[0] == {} and infos[-1]["episode"]["r"] == 2.5
    assert np.array_equal(infos[1]["terminal_observation"], term[0]) and infos[1]["TimeLimit.truncated"] is False
    assert infos[4]["episode"]["l"] == 96 and [bool(i) for i in infos] == [False, True, False, False, True]
    with pytest.raises(IndexError):
        infos[5]


@pytest.mark.parametrize("use_case", ["lmd", "ut", "ct"])
def test_generator_and_builder_invariants(use_case):
    sched = generate_schedule(use_case, 3, start="2020-01-01 00:00", end="2020-02-29 23:59", seed=5)
    assert list(sched.columns) == ["date", "Distance_km", "Consumption_kWh", "Location", "ChargingStation", "ID", "PowerRating_kW"]
    price, tariff, load, pv = synthetic_series(start="2020-01-01 00:00", end="2020-02-29 23:59")
    b = build_fleet(cfgmod.default_config(use_case, time_picker="random", end_cutoff=10), FleetInputs(sched, price, tariff, load, pv))
    tb, c = b.tables, b.consts
    N, T = tb["there"].shape
    assert N == 3 and T == 60 * 96 and c.table_len == T and c.num_evs == 3
    there, tl, sr = tb["there"], tb["time_left"], tb["soc_on_return"]
    assert ((there == 0) | (there == 1)).all()
    assert (tl[there == 0] == 0).all() and (sr[there == 0] == 0).all()
    # while a vehicle stays plugged in, time_left counts down by exactly dt per step
    stay = (there[:, :-1] == 1) & (there[:, 1:] == 1) & (tl[:, 1:] > 0)
    assert np.array_equal(tl[:, :-1][stay] - c.dt, tl[:, 1:][stay])
    # the row before a departure shows exactly one step left
    dep = (there[:, :-1] == 1) & (there[:, 1:] == 0)
    assert (tl[:, :-1][dep] == c.dt).all()
    assert sr.max() <= c.target_soc + 1e-12 and c.start_lo == 0 and 0 < c.start_hi < T
    if use_case == "lmd":   # no Sunday operation
        sunday = np.array([d.astype("datetime64[D]").astype(object).weekday() == 6 for d in b.dates])
        assert (there[:, sunday] == 1).all()
    # reward curves: monthly means equal the global mean (shape_price_reward)
    prc = tb["price_reward_curve"]
    jan = prc[: 31 * 96]
    assert abs(jan.mean() - prc.mean()) < 1e-9


def test_fleet_cache_roundtrip(tmp_path):
    """save_fleet / load_fleet (binary cache of the flattened tables): constants and tables restored bit for bit."""
    from fleetrl_b200.config import default_config
    from fleetrl_b200.schedule import generate_schedule, synthetic_series
    from fleetrl_b200.tables import FleetInputs, build_fleet, load_fleet, save_fleet
    sched = generate_schedule("lmd", 3, start="2020-01-01 00:00", end="2020-02-29 23:59", seed=4)
    price, tariff, load, pv = synthetic_series(start="2020-01-01 00:00", end="2020-02-29 23:59")
    cfg = default_config("lmd", end_cutoff=10, include_pv=False)
    built = build_fleet(cfg, FleetInputs(sched, price, tariff, load, pv))
    path = str(tmp_path / "fleet_cache")
    save_fleet(built, path)
    back = load_fleet(path)
    assert back.consts.to_dict() == built.consts.to_dict()
    assert set(back.tables) == set(built.tables)
    for k, v in built.tables.items():
        if v is None:
            assert back.tables[k] is None
        else:
            assert back.tables[k].dtype == v.dtype
            np.testing.assert_array_equal(back.tables[k], v)
    np.testing.assert_array_equal(back.dates, built.dates)
    assert back.start_ranges == built.start_ranges and back.company == built.company


def test_write_reference_csvs_round_trip(tmp_path):
    """schedule.write_reference_csvs -> tables.read_inputs -> build_fleet gives the same fleet as building it from
    FleetInputs.csv_round_trip() in memory, bit for bit (the unmodified reference reading those files is pinned by the
    *_gen goldens, oracle/gen_golden.py); the raw in-memory frames differ from it by at most one ulp per value."""
    from fleetrl_b200.schedule import write_reference_csvs
    sched = generate_schedule("ct", 3, start="2020-01-01 00:00", end="2020-02-29 23:59", seed=5)
    price, tariff, load, pv = synthetic_series(start="2020-01-01 00:00", end="2020-02-29 23:59", seed=6)
    names = write_reference_csvs(str(tmp_path), "3_ct_gen.csv", sched, price, tariff, load, pv)
    assert set(os.listdir(tmp_path)) == {"3_ct_gen.csv", "synthetic_spot.csv", "synthetic_tariff.csv", "synthetic_load.csv"}
    cfg = cfgmod.default_config("ct", end_cutoff=10, **names)
    from_files = build_fleet(cfg, auto_reset=True)
    raw = FleetInputs(sched, price, tariff, load, pv)
    in_memory = build_fleet(cfg, raw.csv_round_trip(), auto_reset=True)
    assert from_files.consts.to_dict() == in_memory.consts.to_dict()
    for k, v in from_files.tables.items():
        if v is None:
            assert in_memory.tables[k] is None
        else:
            np.testing.assert_array_equal(in_memory.tables[k], v, err_msg=k)
    np.testing.assert_array_equal(from_files.dates, in_memory.dates)
    unrounded = build_fleet(cfg, raw, auto_reset=True)
    for k in ("soc_on_return", "delu", "load"):
        np.testing.assert_allclose(unrounded.tables[k], from_files.tables[k], rtol=1e-13, atol=1e-15, err_msg=k)
    # the schedule file has the reference's schema (data_processing.py:47-61)
    import pandas as pd
    back = pd.read_csv(tmp_path / "3_ct_gen.csv", parse_dates=["date"])
    assert list(back.columns)[1:] == ["date", "Distance_km", "Consumption_kWh", "Location", "ChargingStation", "ID", "PowerRating_kW"]
    assert len(back) == len(sched) and set(back["Location"]) == {"home", "driving"}


def test_gen_schedule_is_honoured(tmp_path):
    """env_config["gen_schedule"]=True (fleet_environment.py:181, auto_gen :969-992): gen_n_evs schedules over
    gen_start_date .. gen_end_date are generated, saved as gen_name in data_path and used instead of schedule_name."""
    from fleetrl_b200.schedule import write_reference_csvs
    sched = generate_schedule("lmd", 1, start="2020-01-01 00:00", end="2020-02-29 23:59", seed=5)
    price, tariff, load, pv = synthetic_series(start="2020-01-01 00:00", end="2020-02-29 23:59", seed=6)
    names = write_reference_csvs(str(tmp_path), "1_lmd.csv", sched, price, tariff, load, pv)
    cfg = cfgmod.default_config("lmd", end_cutoff=10, gen_schedule=True, gen_n_evs=3, gen_name="generated",
                                gen_start_date="2020-01-01 00:00", gen_end_date="2020-02-29 23:59", **names)
    built = build_fleet(cfg, auto_reset=True)
    assert os.path.exists(tmp_path / "generated.csv")
    assert built.consts.num_evs == 3                               # not the 1-EV schedule_name file
    import pandas as pd
    gen = pd.read_csv(tmp_path / "generated.csv")
    assert sorted(gen["ID"].unique()) == [0, 1, 2]
    # in-memory inputs: the generated frame replaces inputs.schedule
    b2 = build_fleet(dict(cfg, gen_n_evs=2), FleetInputs(sched, price, tariff, load, pv), auto_reset=True)
    assert b2.consts.num_evs == 2
    # custom use case: statistics from the custom_* keys (schedule_config.py:134-172)
    c3 = cfgmod.default_config("custom", end_cutoff=10, gen_schedule=True, gen_n_evs=2, gen_start_date="2020-01-01 00:00",
                               gen_end_date="2020-02-29 23:59", custom_ev_battery_size_in_kwh=300, custom_ev_charger_power_in_kw=120,
                               custom_grid_connection_in_kw=500, custom_weekday_distance_mean=200,
                               max_batt_cap_in_all_use_cases=300)
    b3 = build_fleet(c3, FleetInputs(sched, price, tariff, load, pv), auto_reset=True)
    assert b3.consts.num_evs == 2 and b3.consts.evse_max_power == 120.0 and b3.consts.init_battery_cap == 300


def test_unsupported_step_lengths_are_rejected_early():
    with pytest.raises(ValueError, match="float32"):
        cfgmod.resolve(cfgmod.default_config("lmd", freq="5min", minutes=5, time_steps_per_hour=12))
    cfgmod.resolve(cfgmod.default_config("lmd", freq="30min", minutes=30, time_steps_per_hour=2))
