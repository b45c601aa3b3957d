#!/bin/bash
# One GPU-box visit for the round's evidence (B200_PROFILING.md recipe): launch list of a short bench run, one --set full
# capture of the step kernel and of the post kernel in the de-phased steady state, the bench line itself (not under ncu).
tag=${1:-r2}
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --raw-inputs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/launch_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fleet_step_pf -s 150 -c 2 -o gpurun_out/step_$tag -f $B > gpurun_out/ncu_step_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fleet_post -s 150 -c 2 -o gpurun_out/post_$tag -f $B > gpurun_out/ncu_post_$tag.log 2>&1
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
tail -c 600 gpurun_out/bench_${tag}_n1.json
