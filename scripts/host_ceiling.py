"""What limits the end-to-end (host-buffer) step at N GPUs of one box?  Run under torchrun, one rank per GPU.

For the cfg2 shape (65,536 envs x 50 EVs, D = 388: 13.1 MB of actions in, 102 MB of observations out per step and GPU)
every rank measures, with all ranks running at the same time (barrier before each timed loop, max over ranks reported):
  1. plain cudaMemcpyAsync D2H of the observation bytes into page-locked memory      -> the host-side ceiling
  2. plain H2D of the action bytes
  3. fleet_step_host with page-locked buffers used in place by the kernels (zero-copy, the default)
  4. fleet_step_host with staged copies (FLEETSTEP_HOST_ZEROCOPY=0)
  5. 3. with write-combined page-locked buffers (cudaHostAllocWriteCombined)
Results: one JSON line on rank 0 (per-GPU and aggregate GB/s, ms per step, the box's GPU/CPU/NUMA topology)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from fleetrl_b200._lib import FleetStepHandle

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


cudart = C.CDLL("libcudart.so.12")


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    rc = cudart.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc failed: {rc}")
    return p.value


def np_view(ptr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_char * n).from_address(ptr), dtype=dtype).reshape(shape)


sys.argv = [sys.argv[0], "--raw-inputs"]
args = bench.parse_args()
built = bench.build_workload(args)
E, N = args.envs, built.consts.num_evs
out = {"n_gpus": world, "envs_per_gpu": E, "evs": N}

# 1/2: plain copies
D = 388
obs_bytes, act_bytes = E * D * 4, E * N * 4
d_obs = torch.empty(obs_bytes, dtype=torch.uint8, device=dev)
h_obs = torch.empty(obs_bytes, dtype=torch.uint8).pin_memory()
d_act = torch.empty(act_bytes, dtype=torch.uint8, device=dev)
h_act = torch.empty(act_bytes, dtype=torch.uint8).pin_memory()
for name, dst, src, nb in (("d2h_obs", h_obs, d_obs, obs_bytes), ("h2d_act", d_act, h_act, act_bytes)):
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    el = max_over_ranks(time.perf_counter() - t0)
    out[name] = {"ms": el / K * 1e3, "gbs_per_gpu": nb * K / el / 1e9, "gbs_aggregate": world * nb * K / el / 1e9}

# 3-5: the product's host-buffer step
def time_step_host(zero_copy, write_combined):
    os.environ["FLEETSTEP_HOST_ZEROCOPY"] = "1" if zero_copy else "0"
    h = FleetStepHandle(built.consts, built.tables, E, device=lr, env_id_offset=rank * E)
    Dd = h.D
    flags = 0x04 if write_combined else 0x00          # cudaHostAllocWriteCombined
    flags |= 0x02                                       # cudaHostAllocMapped
    an = np_view(host_alloc(E * N * 4, flags), (E, N), np.float32)
    on = np_view(host_alloc(E * Dd * 4, flags), (E, Dd), np.float32)
    rn = np_view(host_alloc(E * 4, 0x02), (E,), np.float32)
    dn = np_view(host_alloc(E, 0x02), (E,), np.uint8)
    an[...] = np.random.default_rng(rank).uniform(-1, 1, (E, N)).astype(np.float32)
    h.reset()
    for _ in range(3):
        h.step_host(an, on, rn, dn)
    barrier()
    K = 10
    t0 = time.perf_counter()
    for _ in range(K):
        h.step_host(an, on, rn, dn)
    el = max_over_ranks(time.perf_counter() - t0)
    h.close()
    nb = E * N * 4 + E * Dd * 4 + E * 5
    return {"ms_per_step": el / K * 1e3, "ev_steps_per_s": world * E * N * K / el, "pcie_gbs_per_gpu": nb * K / el / 1e9,
            "pcie_gbs_aggregate": world * nb * K / el / 1e9}


out["step_host_zero_copy"] = time_step_host(True, False)
out["step_host_staged_copies"] = time_step_host(False, False)
out["step_host_zero_copy_write_combined"] = time_step_host(True, True)
if rank == 0:
    try:
        out["topology"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-3000:]
        out["cpu_affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]
        out["numa_nodes"] = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
    except Exception as e:
        out["topology"] = f"unavailable: {e}"
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
