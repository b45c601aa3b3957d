#!/bin/bash
# usage: gpu_variants.sh "ENV=.. [ARGS=..]" ...   — one bench line per variant (diagnostic runs; ARGS = extra bench.py flags)
show='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("value %.4g  ms/step %.4f  step %.4f  post %.4f  bytes %.3g" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["post_kernel_ms"], d["device_bytes"]))'
for v in "$@"; do
  echo "== $v"
  extra=$(echo "$v" | sed -n 's/.*ARGS=\(.*\)$/\1/p')
  envs=$(echo "$v" | sed 's/ARGS=.*$//')
  env $envs python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --raw-inputs $extra 2>&1 | python -c "$show"
done
