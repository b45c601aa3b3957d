"""Diagnostic: per-step CUDA-event timing of fleet_step at the bench workload (not part of the product)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fleetrl_b200._lib import FleetStepHandle

class A: pass
args = A(); args.use_case="lmd"; args.evs=int(os.environ.get("EVS","50")); args.episode_hours=24; args.carry=int(os.environ.get("CARRY","1")); args.cfg=bench.CONFIGS["cfg2"]; args.raw_inputs=True; args.envs=int(os.environ.get("ENVS","65536"))
built = bench.build_workload(args)
if os.environ.get("NODEG"):
    built.consts.calc_degradation = 0
E, N = args.envs, built.consts.num_evs
dev = torch.device("cuda", 0)
h = FleetStepHandle(built.consts, built.tables, E, device=0)
D = h.D
obs = torch.empty((E, D), dtype=torch.float32, device=dev); term = torch.empty_like(obs)
rew = torch.empty(E, dtype=torch.float32, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
ring = [torch.empty((E, N), dtype=torch.float32, device=dev).uniform_(-1, 1) for _ in range(8)]
h.reset(obs=obs); torch.cuda.synchronize()
stream = torch.cuda.current_stream(dev); sp = stream.cuda_stream
ptrs = [a.data_ptr() for a in ring]
K = 320
evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
h.set_timing(True)
t0 = time.perf_counter()
evs[0].record(stream)
for s in range(K):
    h.step_unchecked(ptrs[s % 8], obs.data_ptr(), rew.data_ptr(), done.data_ptr(), term.data_ptr(), sp)
    evs[s + 1].record(stream)
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
ms = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(K)])
print("cpu enqueue per step us:", t_cpu / K * 1e6)
print("per-step ms: min %.4f median %.4f mean %.4f max %.4f" % (ms.min(), np.median(ms), ms.mean(), ms.max()))
print("slowest steps:", np.argsort(-ms)[:6], np.sort(-ms)[:6] * -1)
print("steps 100..120:", np.round(ms[100:120], 3))
a, b, n = h.get_timing()
print("kernel split (CUDA events inside the library): step kernel %.4f ms, post kernel %.4f ms per step over %d steps" % (a / n, b / n, n))
