#!/bin/bash
# upload policy on THIS box: adaptive raw chunks (capped / uncapped) vs packing everything
out=gpurun_out/r02_sweep11_$1.txt
: > $out
nproc >> $out
run() { echo "## $*" >> $out; env "$@" 2>&1 | grep pinned >> $out; }
for rep in 1 2; do
run ZB_UPLOAD_RAW_CAP=6 python tools/upload_bench.py 29
run ZB_UPLOAD_RAW_CAP=4 python tools/upload_bench.py 29
run ZB_UPLOAD_RAW_CAP=0 python tools/upload_bench.py 29
run ZB_UPLOAD_RAW_EVERY=0 python tools/upload_bench.py 29
done
