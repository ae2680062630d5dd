#!/bin/bash
# Round-2 A/B runs: linear d=1 schedule, bulk-copy rings (one process per setting: the knobs are read once)
out=gpurun_out/r02_sweep1.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
for lg in 20 22 24 28; do
  run ZB_LINEAR_D1=0 python tools/run_case.py sumcheck --log2n $lg --reps 50 --noprofile
  run ZB_LINEAR_D1=1 ZB_HOST_TAIL_LOG2=10 python tools/run_case.py sumcheck --log2n $lg --reps 50 --noprofile
  run ZB_LINEAR_D1=1 ZB_HOST_TAIL_LOG2=5 python tools/run_case.py sumcheck --log2n $lg --reps 50 --noprofile
done
run ZB_LINEAR_D1=1 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_FOLDK_CPS=2 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_FOLDK_CPS=4 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_BSUM_CPS=16 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_EVAL_BULK=0 python tools/run_case.py eval --log2n 28 --reps 20
run ZB_EVAL_BULK=1 python tools/run_case.py eval --log2n 28 --reps 20
run ZB_GRID_BULK=0 python tools/run_case.py prod3 --log2n 30 --reps 10
run ZB_GRID_BULK=4 python tools/run_case.py prod3 --log2n 30 --reps 10
run ZB_GRID_BULK=4 ZB_GRID_BULK_TPB=128 python tools/run_case.py prod3 --log2n 30 --reps 10
run ZB_GRID_BULK=6 python tools/run_case.py prod3 --log2n 30 --reps 10
run ZB_GRID_BULK=6 ZB_GRID_BULK_TPB=128 python tools/run_case.py prod3 --log2n 30 --reps 10
