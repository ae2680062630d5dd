#!/bin/bash
out=gpurun_out/r02_sweep4.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
for u in 8 102 111; do
  run ZB_KECCAK_UNROLL=$u python tools/run_case.py merkle --log2n 26 --reps 5
done
run ZB_KECCAK_UNROLL=102 ZB_KECCAK_UNROLL_LEAF=102 python tools/run_case.py merkle --log2n 26 --reps 5
run ZB_KECCAK_UNROLL=102 python tools/run_case.py merkle --log2n 20 --count 43 --reps 5
run ZB_KECCAK_UNROLL=8 python tools/run_case.py merkle --log2n 20 --count 43 --reps 5
g++ -O2 -std=c++17 -Iinclude tools/c1_trace.cpp -Lzigz_b200 -lzigz_b200 -Wl,-rpath,$PWD/zigz_b200 -o /tmp/c1_trace >> $out 2>&1
run /tmp/c1_trace 20 500
run /tmp/c1_trace 22 200
# ncu: launch list of the bench configuration (only after the same command exited 0 without ncu)
BENCH="python bench.py --steps 2 --warmup 3 --skip-e2e --skip-extras --skip-cpu"
ZB_TAIL_LOG2=0 $BENCH > gpurun_out/r02_bench_prencu.json 2> gpurun_out/r02_bench_prencu.err && \
ZB_TAIL_LOG2=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_prod3_2p30.csv $BENCH > gpurun_out/r02_ncu_list.log 2>&1
ZB_TAIL_LOG2=0 ncu --set full --clock-control none --import-source on -k regex:"k_fold_grid_bulk|k_fold_grid_async|k_round_sums_v4" -c 3 -o gpurun_out/r02_ncu_full_2p30 $BENCH > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_ncu_full_2p30.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_2p30_raw.csv 2>/dev/null
python tools/run_case.py eval --log2n 28 --reps 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_eval_warp10_bulk" -c 1 -o gpurun_out/r02_ncu_eval28 python tools/run_case.py eval --log2n 28 --reps 1 --noprofile > gpurun_out/r02_ncu_eval.log 2>&1
ncu -i gpurun_out/r02_ncu_eval28.ncu-rep --page raw --csv > gpurun_out/r02_ncu_eval28_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_full_2p30.ncu-rep gpurun_out/r02_ncu_eval28.ncu-rep gpurun_out/r02_ncu_lin28.ncu-rep
