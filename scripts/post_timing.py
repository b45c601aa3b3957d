"""Diagnostic: per-phase clock totals of the post kernel (needs a -DPOST_TIMING build copied over libfleetstep.so)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fleetrl_b200._lib import FleetStepHandle, load_library

class A: pass
args = A(); args.use_case="lmd"; args.evs=50; args.episode_hours=24; args.carry=int(os.environ.get("CARRY","1")); args.envs=65536
built = bench.build_workload(args)
E, N = args.envs, built.consts.num_evs
dev = torch.device("cuda", 0)
h = FleetStepHandle(built.consts, built.tables, E, device=0)
L = load_library()
D = h.D
obs = torch.empty((E, D), dtype=torch.float32, device=dev); term = torch.empty_like(obs)
rew = torch.empty(E, dtype=torch.float32, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
ring = [torch.empty((E, N), dtype=torch.float32, device=dev).uniform_(-1, 1) for _ in range(8)]
h.reset(obs=obs); torch.cuda.synchronize()
sp = torch.cuda.current_stream(dev).cuda_stream
buf = (C.c_ulonglong * 16)()
names = ["stage", "phaseA", "phaseB", "stress", "fade", "entry_total", "", "", "nr", "m", "vehicles/32", "new cycles"]
if os.environ.get("FLEETSTEP_POST") == "v1": names[:7] = ["phaseA", "phaseB", "flush", "pass2", "finish", "deg_total", "reset"]
for s in range(200):
    h.step_unchecked(ring[s % 8].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), term.data_ptr(), sp)
    if s in (10, 40, 90, 94, 95, 96, 106, 136, 186, 190):
        L.fleet_debug_post_clk(buf, 1)
        h.step_unchecked(ring[s % 8].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), term.data_ptr(), sp)
        L.fleet_debug_post_clk(buf, 1)
        v = np.array(list(buf), dtype=np.float64)
        w = max(v[10], 1) if os.environ.get('FLEETSTEP_POST') == 'v1' else 8 * 683
        print(f"step {s+1}: warps={int(v[10])} avg cycles/warp: " + " ".join(f"{names[k]}={v[k]/w:.0f}" for k in range(7)) +
              f" | nr/warp-lane0={v[8]/w:.1f} m={v[9]/w:.1f} newcyc={v[11]/w:.2f}")
