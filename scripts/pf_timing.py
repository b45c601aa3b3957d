"""Diagnostic: where the persistent step kernel's warps spend their cycles (needs a -DPF_TIMING build)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fleetrl_b200._lib import FleetStepHandle, load_library

sys.argv = [sys.argv[0], "--raw-inputs", "--config", os.environ.get("CONFIG", "cfg2")]
args = bench.parse_args()
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
for s in range(30):
    h.step_unchecked(ring[s % 8].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), term.data_ptr(), sp)
L.fleet_debug_pf_clk(buf, 1)
K = 20
for s in range(K):
    h.step_unchecked(ring[s % 8].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), term.data_ptr(), sp)
L.fleet_debug_pf_clk(buf, 1)
v = np.array(list(buf), dtype=np.float64)
B = 256 // N
ntiles = (E + B - 1) // B
print("epilogue detail: any_reset %.0f | bulk store issue %.0f" % (v[7] / (K * ntiles), v[6] / (K * ntiles)))
ep = v[:6] / (K * ntiles)
print("epilogue warp, cycles per tile: pre-Done %.0f | wait Done %.0f | store+sums %.0f | wait_read %.0f | stage_env+arrive %.0f | finalise %.0f | total %.0f"
      % (*ep, ep.sum()))
cw = v[[12, 8, 10, 9, 13, 14, 11]] / (K * ntiles * 8)
print("compute warps, cycles per tile: loop top + wait own copies %.0f | wait Env %.0f | wait Free %.0f | body (math, smem stores) %.0f | proxy fence + arrive %.0f | state stores %.0f | issue copies %.0f | total %.0f" % (*cw, cw.sum()))
print("epilogue warps by (%warpid & 3), summed over the launches:", v[12:16] / K)
