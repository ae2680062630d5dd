#!/bin/bash
out=gpurun_out/r02_sweep3.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
g++ -O2 -std=c++17 -Iinclude tools/c1_trace.cpp -Lzigz_b200 -lzigz_b200 -Wl,-rpath,$PWD/zigz_b200 -o /tmp/c1_trace >> $out 2>&1
run /tmp/c1_trace 20 500
run /tmp/c1_trace 22 200
run ZB_HOST_TAIL_LOG2=5 /tmp/c1_trace 20 500
run ZB_LINEAR_D1=0 /tmp/c1_trace 20 500
run ZB_FOLDK_CPS=1 python tools/run_case.py sumcheck --log2n 28 --reps 10
run ZB_FOLDK_CPS=2 python tools/run_case.py sumcheck --log2n 28 --reps 10
python tools/run_case.py sumcheck --log2n 28 --reps 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_foldk_sums|k_block_sums" -c 4 -o gpurun_out/r02_ncu_lin28 python tools/run_case.py sumcheck --log2n 28 --reps 1 --noprofile > gpurun_out/r02_ncu_lin28.log 2>&1
ncu -i gpurun_out/r02_ncu_lin28.ncu-rep --page raw --csv > gpurun_out/r02_ncu_lin28_raw.csv 2>/dev/null
