#!/bin/bash
# One GPU-box visit: parity tests, bench, optional library variants (args: FLEETSTEP_LIB=... settings)
python -m pytest tests -m gpu -q -x 2>&1 | tail -12
bash scripts/gpu_variants.sh "A=1" "$@"
if [ -f fleetrl_b200/libfleetstep_timing.so ]; then FLEETSTEP_LIB=fleetrl_b200/libfleetstep_timing.so python scripts/post_timing.py 2>&1 | tail -14; fi
