#!/bin/bash
# One GPU-box visit: parity tests, bench, optional library variants (args: FLEETSTEP_LIB=... settings)
python -m pytest tests -m gpu -q -x 2>&1 | tail -12
bash scripts/gpu_variants.sh "A=1" "$@"
