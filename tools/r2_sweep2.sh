#!/bin/bash
# Round-2 A/B runs, second pass: grid-stride linear kernels, K sweep, eval variants
out=gpurun_out/r02_sweep2.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
for lg in 20 22 28; do
  run ZB_LINEAR_D1=1 python tools/run_case.py sumcheck --log2n $lg --reps 50 --noprofile
  run ZB_LINEAR_D1=1 ZB_HOST_TAIL_LOG2=5 python tools/run_case.py sumcheck --log2n $lg --reps 50 --noprofile
  run ZB_LINEAR_K=4 python tools/run_case.py sumcheck --log2n $lg --reps 50 --noprofile
done
run ZB_LINEAR_K=5 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_LINEAR_K=4 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_LINEAR_K=3 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_FOLDK_CPS=2 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_FOLDK_CPS=4 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_BSUM_CPS=16 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_BSUM_CPS=4 python tools/run_case.py sumcheck --log2n 28 --reps 10
for b in 0 1 2 3 4 5 6; do
  run ZB_EVAL_BULK=$b python tools/run_case.py eval --log2n 28 --reps 20
done
run ZB_EVAL_BULK=0 ZB_EVAL_UT=4 python tools/run_case.py eval --log2n 28 --reps 20
run ZB_EVAL_BULK=0 ZB_EVAL_CPS=4 python tools/run_case.py eval --log2n 28 --reps 20
run ZB_EVAL_BULK=0 python tools/run_case.py eval --log2n 20 --reps 50 --noprofile
run ZB_EVAL_BULK=0 python tools/run_case.py eval --log2n 24 --reps 50 --noprofile
