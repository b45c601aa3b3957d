"""Print the headline metrics of an .ncu-rep (one kernel per row): python profiles/ncu_metrics.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.max', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block', 'sm__throughput.avg.pct',
        'smsp__inst_executed_pipe_fp64', 'sm__pipe_fp64_cycles_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'smsp__issue_active.avg.pct', 'lts__t_bytes.sum ', 'sm__inst_executed_pipe_lsu', 'smsp__inst_executed_op_shared']
skip = ['per_second', 'pct_of_peak_sustained_elapsed', '.peak_sustained']
for v in rows[2:]:
    print('=' * 100)
    for i, n in enumerate(h):
        if any(w in n for w in want) and (not any(k in n for k in skip) or 'dram__throughput' in n):
            if v[i] not in ('', '0'):
                print(f"{n:90s} {u[i]:14s} {v[i]}")
