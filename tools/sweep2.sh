#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_cpp_api.py -m gpu -q 2>&1 | tail -2
for u in 1 2 4; do for cps in 2 4 8; do echo "== d1 fold U=$u CPS=$cps"; ZB_FOLD_U=$u ZB_FOLD_CPS=$cps python tools/fold_bench.py 29 | head -1; done; done
for cps in 4 8 16 32; do echo "== rsum CPS=$cps"; ZB_RSUM_CPS=$cps python tools/fold_bench.py 29 | sed -n 2,3p; done
