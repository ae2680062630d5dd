#!/bin/bash
# tuning sweep (round 1): Keccak FMA-rotation share, eval launch shape, d=1 fold unroll
cd "$(dirname "$0")/.."
echo "== pytest gpu (default variant)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 2 3 4; do echo "== merkle ZB_KECCAK_V=$v"; ZB_KECCAK_V=$v python tools/run_case.py merkle --log2n 24 --reps 5 | head -4; done
echo "== merkle parity with V=4"; ZB_KECCAK_V=4 timeout 600 python -m pytest tests/test_gpu_merkle.py -x -q 2>&1 | tail -2
echo "== merkle parity with V=2"; ZB_KECCAK_V=2 timeout 600 python -m pytest tests/test_gpu_merkle.py -x -q 2>&1 | tail -2
for ut in 1 2 4; do for cps in 2 4 8; do echo "== eval UT=$ut CPS=$cps"; ZB_EVAL_UT=$ut ZB_EVAL_CPS=$cps python tools/run_case.py eval --log2n 28 --reps 10 | head -3; done; done
for u in 1 2 4; do for cps in 2 4 8; do echo "== d1 fold U=$u CPS=$cps"; ZB_FOLD_U=$u ZB_FOLD_CPS=$cps python tools/run_case.py sumcheck --log2n 28 --reps 5 | head -3; done; done
for cps in 4 8 16; do echo "== rsum CPS=$cps"; ZB_RSUM_CPS=$cps python tools/run_case.py sumcheck --log2n 28 --reps 5 | sed -n 3p; done
for cps in 2 4 8; do echo "== d3 fold CPS=$cps"; ZB_FOLD_CPS=$cps python tools/run_case.py prod3 --log2n 28 --reps 5 | head -3; done
