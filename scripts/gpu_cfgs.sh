#!/bin/bash
# One-GPU visit: the BASELINE.json configurations as bench lines (cfg2 is the metric's; cfg3 / cfg4 / cfg4full / cfg5 and
# the uncontrolled-charging run of cfg2 are reported beside it), then the reference arm.
for c in cfg3 cfg4 cfg4full cfg5; do
  python bench.py --config $c --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r2_${c}_n1.json 2> gpurun_out/bench_r2_${c}_n1.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_r2_${c}_n1.json').read().strip().split('\n')[-1]);print('$c',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['roofline']['post_kernel_ms'],d['e2e']['value'])"
done
python bench.py --config cfg2 --actions uncontrolled --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2_cfg2_uncontrolled_n1.json 2>&1
tail -c 300 gpurun_out/bench_r2_cfg2_uncontrolled_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2>&1
tail -c 600 gpurun_out/bench_r2_reference.json
