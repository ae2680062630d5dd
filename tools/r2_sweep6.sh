#!/bin/bash
out=gpurun_out/r02_sweep6.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
run ZB_GRID_FIRST=sums python tools/run_case.py prod3 --log2n 30 --reps 10
run ZB_GRID_FIRST=grid python tools/run_case.py prod3 --log2n 30 --reps 10
run ZB_GRID_FIRST=grid ZB_GRID_BULK=5 python tools/run_case.py prod3 --log2n 30 --reps 10
run ZB_GRID_FIRST=grid ZB_GRID_ASYNC=0 python tools/run_case.py prod3 --log2n 30 --reps 10
