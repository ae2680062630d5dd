#!/bin/bash
out=gpurun_out/r02_sweep5.txt
: > $out
run() { echo "## $*" >> $out; env "$@" >> $out 2>&1; }
python -m pytest tests/test_gpu_prover_glue.py tests/test_gpu_merkle.py -m gpu -x -q >> $out 2>&1
for b in 18 22 24 26 40; do
  run ZB_GRID_BULK_MIN_LOG2=$b python tools/run_case.py prod3 --log2n 30 --reps 10
done
run ZB_GRID_BULK_MIN_LOG2=18 python tools/run_case.py prod3 --log2n 27 --reps 20
run ZB_GRID_BULK_MIN_LOG2=40 python tools/run_case.py prod3 --log2n 27 --reps 20
python bench.py --log2n 28 --cpu-log2n 22 --skip-e2e --skip-cpu > gpurun_out/r02_bench28_check3.json 2> gpurun_out/r02_bench28_check3.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench28_check3.json'))
for k in ('C4_witness_43x2^20','C4_prove_from_trace_2^18_steps'):
    v=d['extras'][k]; v.pop('cpu_baseline',None); v.pop('note',None); print(k, json.dumps(v)[:600])
" >> $out 2>&1
