#!/bin/bash
# One GPU-box visit: parity tests, two bench variants (history ring 32 / 64 rows) and an ncu capture of the post kernel.
python -m pytest tests -m gpu -q -x 2>&1 | tail -12
show='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("value %.4g  ms/step %.4f  step %.4f  post %.4f  bytes %.3g" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["post_kernel_ms"], d["device_bytes"]))'
python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline 2>&1 | python -c "$show"
FLEETSTEP_RF_RING=64 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline 2>&1 | python -c "$show"
ncu --set full --clock-control none --import-source on -k regex:fleet_post_kernel -s 150 -c 1 -o gpurun_out/post_r2e python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_post.log 2>&1
tail -2 gpurun_out/ncu_post.log | cut -c1-200
