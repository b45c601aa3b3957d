"""cfg5 of SURVEY §8(d): a PPO-style rollout through the SB3-shaped FleetVecEnv, sharded over the GPUs of one node.

  python scripts/rollout_cfg5.py --total-envs 1048576 --steps 64
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
      scripts/rollout_cfg5.py --total-envs 1048576 --steps 64

Every rank owns total_envs / G environments (FleetVecEnv.sharded), runs a 64x64 tanh MLP policy on its own GPU (random
weights, Gaussian action sampling clipped to [-1, 1]: the shape of SB3's MlpPolicy), steps the fleet with
`step_raw` (no host round trip) and, once per rollout, all-reduces the episode statistics over NCCL.  It prints ONE JSON
line: env-only EV-steps/s (CUDA events around the env steps, policy excluded) and end-to-end rollout EV-steps/s
(policy + env), both as the max-over-ranks time.  This is a measurement script, not part of the product path."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from fleetrl_b200 import FleetVecEnv
from fleetrl_b200.dist import shard_range, world


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-envs", type=int, default=1048576)
    ap.add_argument("--evs", type=int, default=50)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    a = ap.parse_args()
    rank, ws, local = world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)

    class W: pass
    w = W(); w.use_case = "lmd"; w.evs = a.evs; w.episode_hours = 24; w.carry = 1; w.cfg = bench.CONFIGS["cfg5"]; w.raw_inputs = True
    built = bench.build_workload(w)
    lo, hi = shard_range(a.total_envs, rank, ws)
    env = FleetVecEnv(None, hi - lo, device=local, env_id_offset=lo, built=built, output="torch")
    E, N, D = env.num_envs, env.num_cars, env.obs_dim

    g = torch.Generator(device=dev); g.manual_seed(1234)      # same policy weights on every rank
    def lin(i, o):
        return (torch.randn(i, o, device=dev, generator=g) / i ** 0.5).to(torch.bfloat16), torch.zeros(o, device=dev, dtype=torch.bfloat16)
    w1, b1 = lin(D, 64); w2, b2 = lin(64, 64); w3, b3 = lin(64, N)
    log_std = -0.5

    def policy(obs):
        h = torch.tanh(torch.addmm(b1, obs.to(torch.bfloat16), w1))
        h = torch.tanh(torch.addmm(b2, h, w2))
        mean = torch.addmm(b3, h, w3).float()
        return (mean + torch.randn_like(mean) * (2.718281828 ** log_std)).clamp_(-1, 1)

    obs = env.reset()
    stream = torch.cuda.current_stream(dev)
    # de-phase the episodes like bench.py does (env e is e mod L steps into its episode): every rollout step then sees its
    # share of auto-resets and daily evaluations, the SB3 steady state
    L = int(built.consts.episode_steps)
    phase = torch.arange(E, device=dev, dtype=torch.int64) % L
    for s in range(L):
        obs, _, _ = env.step_raw(policy(obs))
        env.handle.reset(mask=(phase == s).to(torch.uint8), obs=env._obs)
    for _ in range(a.warmup):
        obs, _, _ = env.step_raw(policy(obs))
    torch.cuda.synchronize(dev)
    env.handle.reset_stats()
    if ws > 1:
        dist.barrier()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for s in range(a.steps):
        act = policy(obs)
        e0[s].record(stream)
        obs, rew, done = env.step_raw(act)
        e1[s].record(stream)
    stats = env.handle.stats_tensor()
    if ws > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)          # once per rollout
    t1.record(stream)
    torch.cuda.synchronize(dev)
    env_ms = sum(x.elapsed_time(y) for x, y in zip(e0, e1))
    tot_ms = t0.elapsed_time(t1)
    tm = torch.tensor([env_ms, tot_ms], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    env_ms, tot_ms = tm.tolist()
    if rank == 0:
        ev_steps = a.total_envs * N * a.steps
        print(json.dumps({
            "config": f"cfg5: {a.total_envs} envs x {N} EVs over {ws} GPU(s), PPO-style rollout of {a.steps} steps, MLP 64x64 policy on device",
            "n_gpus": ws, "envs_per_gpu": E, "obs_dim": D,
            "env_only_ev_steps_per_s": ev_steps / (env_ms * 1e-3), "env_ms_per_step": env_ms / a.steps,
            "rollout_ev_steps_per_s": ev_steps / (tot_ms * 1e-3), "rollout_ms_per_step": tot_ms / a.steps,
            "episodes_finished": float(stats[0].item()), "env_steps": float(stats[2].item()),
            "device_bytes_per_gpu": env.handle.device_bytes}), flush=True)
    env.close()
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
