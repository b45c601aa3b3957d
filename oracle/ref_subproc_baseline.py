"""CPU baseline of BASELINE.md §3 / SURVEY §8d: the UNMODIFIED reference FleetEnv stepped in a SubprocVecEnv-style harness
(one worker process per env, pipe round trip per step, reset on done — stable-baselines3's protocol; SB3 itself is not
installable here), on all cores available to this process.

TEST / MEASUREMENT INFRASTRUCTURE, build container only (needs /root/reference).  Usage:
    python oracle/ref_subproc_baseline.py cfg1            # shipped 1-EV lmd schedule, one 96-step episode per worker
    python oracle/ref_subproc_baseline.py cfg2 [steps]    # 50-EV lmd fleet from the product's generator (written to the
                                                          #  reference's CSV dialects), `steps` steps per worker (default 24)
Prints one JSON line: aggregate EV-steps/s, ms per env-step, cores."""
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, os.path.dirname(HERE))


def worker(conn, cfg, seed):
    import numpy as np
    import compat
    env = compat.make_reference_env(cfg)
    rng = np.random.default_rng(seed)
    conn.send(("ready", int(env.num_cars)))
    while True:
        cmd, arg = conn.recv()
        if cmd == "reset":
            conn.send(env.reset()[0])
        elif cmd == "step":
            obs, r, done, trunc, info = env.step(arg)
            if done:                                   # SubprocVecEnv worker: reset on done, return the new observation
                info = dict(info, terminal_observation=obs)
                obs = env.reset()[0]
            conn.send((obs, r, done, info))
        else:
            conn.close()
            return


def main():
    import numpy as np
    import compat
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    cores = len(os.sched_getaffinity(0))
    if which == "cfg1":
        cfg = compat.base_config(spot_markup=0, spot_mul=1, feed_in_ded=0, time_picker="static")
        steps = 96
    else:
        import gen_golden
        steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
        cfg = compat.base_config(time_picker="static", price_name=None, tariff_name=None)
        cfg.update(gen_golden.make_generated_data_path("lmd", 50, "subproc_cfg2", 101))
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for k in range(cores):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=worker, args=(b, cfg, k), daemon=True)
        pr.start()
        pipes.append(a); procs.append(pr)
    n_cars = [p.recv()[1] for p in pipes][0]
    for p in pipes:
        p.send(("reset", None))
    for p in pipes:
        p.recv()
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    for s in range(steps):
        acts = rng.uniform(-1, 1, (cores, n_cars))
        for p, a in zip(pipes, acts):                  # step_async
            p.send(("step", a))
        for p in pipes:                                # step_wait
            p.recv()
    el = time.perf_counter() - t0
    for p in pipes:
        p.send(("close", None))
    print(json.dumps({"config": which, "reference": "unmodified /root/reference FleetEnv, SubprocVecEnv-style harness",
                      "cores": cores, "n_envs": cores, "num_cars": n_cars, "steps_per_env": steps, "seconds": el,
                      "ev_steps_per_s": cores * n_cars * steps / el, "ms_per_env_step": el / steps * 1e3}))


if __name__ == "__main__":
    main()
