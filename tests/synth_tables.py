"""Small synthetic canonical tables for parity tests (independent of the product's table builder).

Produces the FleetTables arrays directly: per-vehicle presence patterns with one or two trips per day, time_left /
SOC_on_return derived with the same definitions as the reference DataLoader (data_processing.py:120-223), hourly
price / tariff / load / pv series held constant within the hour, monthly de-trended reward curves replaced by a
simple smooth offset (any values are valid inputs for the step).
"""
import numpy as np

from fleetrl_b200._abi import FleetConsts


def make_tables(n_evs=7, days=12, sph=4, seed=0, two_trips=False, start_weekday=2, cap=60.0, target_soc=0.85):
    rng = np.random.default_rng(seed)
    spd = 24 * sph
    T = days * spd
    dt = 1.0 / sph
    there = np.ones((n_evs, T), np.uint8)
    cons_ret = np.zeros((n_evs, T))           # last trip consumption attached from the return row on
    for n in range(n_evs):
        last_cons = 0.0
        for d in range(days):
            wd = (start_weekday + d) % 7
            if wd == 6 and not two_trips:
                continue
            trips = []
            dep = int(np.clip(rng.normal(7, 1), 3, 11) * sph)
            ret = int(np.clip(rng.normal(19, 1), 12, 23) * sph)
            if two_trips:
                pb = int(np.clip(rng.normal(12, 0.25), 11.5, 12.75) * sph)
                pe = int(np.clip(rng.normal(13.5, 0.25), 13, 14.5) * sph)
                trips = [(dep, pb), (pe, ret)]
            else:
                trips = [(dep, ret)]
            for (a, b) in trips:
                if b <= a:
                    continue
                there[n, d * spd + a:d * spd + b] = 0
                c = float(np.clip(rng.normal(25, 12), 2, 0.8 * cap))
                cons_ret[n, d * spd + b:] = c
    # time_left: hours until the next departure row (first absent row), 0 when absent or no departure ahead
    time_left = np.zeros((n_evs, T))
    for n in range(n_evs):
        nxt = -1
        for t in range(T - 1, -1, -1):
            if there[n, t] == 0:
                if t == 0 or there[n, t - 1] == 1:
                    nxt = t  # departure row
                time_left[n, t] = 0
            else:
                time_left[n, t] = (nxt - t) * dt if nxt >= 0 else 0.0
    soc_on_return = np.where(there == 1, target_soc - cons_ret / cap, 0.0)
    hours = T // sph
    delu_h = 40 + 25 * np.sin(np.arange(hours) * 2 * np.pi / 24) + rng.normal(0, 8, hours)
    delu_h[rng.random(hours) < 0.02] *= -0.5                                   # occasional negative prices
    tariff_h = np.round(delu_h * 0.9 + 3, 2)
    load_h = 35 + 30 * np.clip(np.sin((np.arange(hours) % 24 - 6) * np.pi / 12), 0, None) + rng.normal(0, 2, hours)
    pv_h = 60 * np.clip(np.sin((np.arange(hours) % 24 - 6) * np.pi / 12), 0, None) * rng.uniform(0.2, 1, hours)
    rep = lambda x: np.repeat(x, sph)[:T]
    step = np.arange(T)
    hour = ((step // sph) % 24).astype(np.uint8)
    minute = ((step % sph) * (60 // sph)).astype(np.uint8)
    day = step // spd
    weekday = (start_weekday + day) % 7
    month = 1 + (day // 30) % 12
    cal = np.stack([np.sin(2 * np.pi * month / 12), np.cos(2 * np.pi * month / 12),
                    np.sin(2 * np.pi * weekday / 7), np.cos(2 * np.pi * weekday / 7),
                    np.sin(2 * np.pi * hour / 24), np.cos(2 * np.pi * hour / 24)], axis=1)
    delu = rep(delu_h)
    tariff = rep(tariff_h)
    tables = dict(there=there, time_left=time_left, soc_on_return=soc_on_return, delu=delu, tariff=tariff,
                  load=rep(load_h), pv=rep(pv_h),
                  price_reward_curve=(delu + 10) * 1.5 - 3.0 * np.sin(day / 9.0),
                  tariff_reward_curve=tariff * 0.75 + 2.0 * np.cos(day / 7.0),
                  cal_sincos=cal, hour=hour, minute=minute)
    return tables, T


def make_consts(tables, T, n_evs, sph=4, episode_hours=24, use_case="lmd", **over):
    uc = dict(lmd=(11.0, 60.0, 60.0), ut=(22.0, 50.0, 50.0), ct=(4.6, 16.7, 16.7))[use_case]
    evse, lc_cap, cap0 = uc
    max_load = float(tables["load"].max()) if tables.get("load") is not None else 0.0
    grid = max(max_load * 1.1, max_load + 0.5 * n_evs * evse)
    if use_case == "ut" and n_evs > 1:
        grid = 1000.0
    d = dict(
        num_evs=n_evs, table_len=T, steps_per_hour=sph, episode_steps=episode_hours * sph, price_lookahead=8,
        bl_pv_lookahead=4, include_price=1, include_building=1, include_pv=1, aux=1, normalize=0,
        is_caretaker=int(use_case == "ct"), calc_degradation=1, deg_mode=0, carry_degradation_state=1, auto_reset=1,
        start_lo=0, start_hi=T - (episode_hours + 12) * sph, seed=1234,
        dt=1.0 / sph, init_battery_cap=cap0, obc_max_power=100.0, charging_eff=0.91, discharging_eff=0.91,
        def_soc=0.5, temperature=25.0, target_soc=0.85, target_soc_lunch=0.65, min_laxity=2.0, fixed_markup=10.0,
        variable_multiplier=1.5, feed_in_deduction=0.25, evse_max_power=evse, grid_connection=grid,
        lc_batt_cap=lc_cap, price_multiplier=3.33 * 60.0 / cap0, fully_charged_reward=1.0,
        penalty_invalid_action=-0.2, penalty_overcharging=-0.0055, penalty_overloading=1.0, clip_overcharging=-0.2,
        init_soh=1.0, soc_eps=0.005,
        max_time_left=float(tables["time_left"].max()),
        min_price=(float(tables["delu"].min()) + 10.0) * 1.5, max_price=(float(tables["delu"].max()) + 10.0) * 1.5,
        min_tariff=float(tables["tariff"].min()) * 0.75, max_tariff=float(tables["tariff"].max()) * 0.75,
        max_building=max_load, max_pv=float(tables["pv"].max()) if tables.get("pv") is not None else 0.0,
    )
    d.update(over)
    return FleetConsts.from_dict(d)
