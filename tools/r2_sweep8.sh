#!/bin/bash
# product-prover latency: small tables finish on the host (ZB_PROD_HOST_TAIL_LOG2) vs one host round trip per round
out=gpurun_out/r02_sweep8_$1.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
for ht in 0 9 10 11 12; do
  for lg in 10 14 20 24; do
    run ZB_PROD_HOST_TAIL_LOG2=$ht python tools/run_case.py prod3 --log2n $lg --reps 300 --noprofile
  done
done
run ZB_PROD_HOST_TAIL_LOG2=0 python tools/run_case.py prod3 --log2n 30 --reps 10 --noprofile
run ZB_PROD_HOST_TAIL_LOG2=9 python tools/run_case.py prod3 --log2n 30 --reps 10 --noprofile
run ZB_PROD_HOST_TAIL_LOG2=9 ZB_GRID_MIN_LOG2=15 python tools/run_case.py prod3 --log2n 20 --reps 300 --noprofile
run ZB_PROD_HOST_TAIL_LOG2=9 ZB_GRID_MIN_LOG2=11 python tools/run_case.py prod3 --log2n 20 --reps 300 --noprofile
