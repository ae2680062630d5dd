#!/bin/bash
out=gpurun_out/r02_sweep7_$1.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
g++ -O2 -std=c++17 -Iinclude tools/c1_trace.cpp -Lzigz_b200 -lzigz_b200 -Wl,-rpath,$PWD/zigz_b200 -o /tmp/c1_trace >> $out 2>&1
run /tmp/c1_trace 20 1000
run python tools/run_case.py prod3 --log2n 20 --reps 300 --noprofile
run python tools/run_case.py prod3 --log2n 14 --reps 300 --noprofile
run python tools/run_case.py prod3 --log2n 30 --reps 10 --noprofile
