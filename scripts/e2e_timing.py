"""Diagnostic: fleet_step_host (pinned host buffers) per-step wall time at the bench workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fleetrl_b200._lib import FleetStepHandle

class A: pass
args = A(); args.use_case="lmd"; args.evs=50; args.episode_hours=24; args.carry=1; args.cfg=bench.CONFIGS["cfg2"]; args.raw_inputs=True; args.envs=65536
built = bench.build_workload(args)
E, N = args.envs, built.consts.num_evs
h = FleetStepHandle(built.consts, built.tables, E, device=0)
D = h.D
a_host = [torch.empty((E, N), dtype=torch.float32).uniform_(-1, 1).pin_memory() for _ in range(2)]
o_host = torch.empty((E, D), dtype=torch.float32).pin_memory()
r_host = torch.empty(E, dtype=torch.float32).pin_memory(); d_host = torch.empty(E, dtype=torch.uint8).pin_memory()
obs = torch.empty((E, D), dtype=torch.float32, device="cuda")
h.reset(obs=obs); torch.cuda.synchronize()
an = [a.numpy() for a in a_host]; on, rn, dn = o_host.numpy(), r_host.numpy(), d_host.numpy()
for s in range(5):
    h.step_host(an[s % 2], on, rn, dn)
ts = []
for s in range(40):
    t0 = time.perf_counter(); h.step_host(an[s % 2], on, rn, dn); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("fleet_step_host ms per step: min %.3f median %.3f mean %.3f  -> %.3e EV-steps/s" % (ts.min(), np.median(ts), ts.mean(), E * N / (ts.mean() * 1e-3)))
print("obs checksum", float(on.sum()))
