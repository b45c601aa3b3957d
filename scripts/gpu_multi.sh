#!/bin/bash
# One multi-GPU box visit (gpurun --gpus N): host ceiling, bench cfg2 e2e + cfg5, cfg5 rollout.  Usage: gpu_multi.sh N
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR scripts/host_ceiling.py > gpurun_out/host_ceiling_n$N.json 2> gpurun_out/host_ceiling_n$N.err
tail -c 1500 gpurun_out/host_ceiling_n$N.json
$TR bench.py --gpus $N --config cfg5 --steps 200 --warmup 20 --no-cpu-baseline --raw-inputs > gpurun_out/bench_r2_cfg5_n$N.json 2> gpurun_out/bench_r2_cfg5_n$N.err
tail -c 400 gpurun_out/bench_r2_cfg5_n$N.json
$TR scripts/rollout_cfg5.py > gpurun_out/cfg5_rollout_n$N.json 2> gpurun_out/cfg5_rollout_n$N.err
tail -c 600 gpurun_out/cfg5_rollout_n$N.json
