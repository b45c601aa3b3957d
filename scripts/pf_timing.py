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
buf = (C.c_ulonglong * 144)()
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

print("one extra ld.volatile.shared round trip in the compute warps (PF_TIMING_LDS builds): %.0f cycles" % (v[15] / (K * ntiles * 8)))
print("per compute warp, cycles per tile: [wait Env | body | wait Free | issue copies | wait own copies | fence+arrive | state stores]  total")
for w in range(8):
    r = v[16 + w * 8: 16 + w * 8 + 8] / (K * ntiles)
    print("  warp %d: Env %5.0f  body %5.0f  Free %5.0f  issue %4.0f  copies %5.0f  fence %4.0f  stores %4.0f   total %5.0f" % (w, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[:7].sum()))

tr = (C.c_ulonglong * 640)()
L.fleet_debug_pf_trace(tr)
t = np.array(list(tr), dtype=np.float64).reshape(10, 8, 8)
t0 = t[t > 0].min()
names_c = ["copies landed", "env ok", "free ok", "arrived done", "copies issued"]
print("timeline of CTA 0 (cycles since the first event), iterations 16..21")
for it in range(6):
    print(" it %d" % (16 + it))
    for w in (0, 3, 7):
        print("   compute warp %d: " % w + "  ".join("%s %6.0f" % (names_c[e], t[w, it, e] - t0) for e in range(5)))
    print("   epilogue 1: " + "  ".join("%s %6.0f" % (n, t[9, it, e] - t0) for e, n in enumerate(["done seen", "store issued", "sums_free ok", "sums ready", "freed"])))
    print("   epilogue 0: " + "  ".join("%s %6.0f" % (n, t[8, it, e] - t0) for e, n in enumerate(["env(it+3) staged", "sums seen", "finalised"])))
