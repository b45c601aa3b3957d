#!/usr/bin/env python
"""Turn one gpurun evidence visit (scripts/gpu_profile.sh <tag>) into the tracked summaries under profiles/:
   launches_<tag>_summary.csv   per-kernel launch count / avg / min / max / share from the ncu launch list
   ncu_<tag>_summary.md         headline --set full metrics + top stall reasons of the step and the post kernel
   step_kernel_traffic.json     dram bytes per launch of the step kernel (read by bench.py for roofline.traffic)
usage: python profiles/summarize.py <tag> [gpurun_out]"""
import collections, csv, json, os, re, subprocess, sys

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
here = os.path.dirname(os.path.abspath(__file__))


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(src, f"launches_{tag}.csv"), errors="replace")) if len(r) > 10]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
        name = re.sub(r"\(.*$", "", r[ik]).replace("<unnamed>::", "").replace("void ", "")[:90]
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = os.path.join(here, f"launches_{tag}_summary.csv")
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,avg_us,min_us,max_us,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k}\",{len(v)},{sum(v):.1f},{sum(v) / len(v):.2f},{min(v):.2f},{max(v):.2f},{100 * sum(v) / tot:.1f}\n")
    print(open(out).read())


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(rep, md):
    out = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    first = None
    for v in rows[2:3]:                                   # first captured launch
        d = dict(zip(h, v))
        first = d
        md.write(f"\n## {d['Kernel Name'][:80]}  ({rep})\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in WANT:
            if k in d and d[k] != "":
                md.write(f"| {k} | {d[k]} | {u[h.index(k)]} |\n")
        stalls = []
        for k in h:
            m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", k)
            if m and not m.group(1).endswith("not_issued"):
                try:
                    stalls.append((float(d[k]), m.group(1)))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1
        md.write("\nWarp-state samples (all): " + ", ".join(f"{n} {100 * s / tot:.1f} %" for s, n in sorted(stalls, reverse=True)[:8]) + "\n")
    return first


if __name__ == "__main__":
    launch_list()
    with open(os.path.join(here, f"ncu_{tag}_summary.md"), "w") as md:
        md.write(f"# ncu --set full summaries, {tag}\n\n`scripts/gpu_profile.sh {tag}` on one B200 (cfg2: 65,536 envs x 50 EVs, de-phased steady "
                 "state; `--clock-control none`, launch 150 onwards).  Times under ncu are cold-cache and serialised; the bench "
                 "line uses live CUDA-event times.\n")
        st = full(f"step_{tag}.ncu-rep", md)
        full(f"post_{tag}.ncu-rep", md)
    rd, wr = float(st["dram__bytes_read.sum"]), float(st["dram__bytes_write.sum"])
    ui = {k: i for i, k in enumerate([])}
    # units of the dram byte counters are reported per metric; normalise to bytes
    out = subprocess.run(["ncu", "-i", os.path.join(src, f"step_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale[u[h.index("dram__bytes_read.sum")]]; wr *= scale[u[h.index("dram__bytes_write.sum")]]
    bench = json.loads(open(os.path.join(src, f"bench_{tag}_n1.json")).read().strip().split("\n")[-1])
    cfg = bench["config"]
    json.dump({"kernel": st["Kernel Name"][:60], "envs": cfg["envs_per_gpu"], "evs": cfg["evs"], "obs_dim": cfg["obs_dim"],
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "source": f"profiles/ncu_{tag}_summary.md (ncu --set full, {tag})"},
              open(os.path.join(here, "step_kernel_traffic.json"), "w"), indent=1)
    print(open(os.path.join(here, f"ncu_{tag}_summary.md")).read())
