#!/bin/bash
# upload staging: chunk size x store type (can the staging buffers live in the LLC?)
out=gpurun_out/r02_sweep9_$1.txt
: > $out
run() { echo "## $*" >> $out; env "$@" 2>&1 | grep pinned >> $out; }
for nt in 1 0; do
  for lg in 23 21 20 19 18; do
    run ZB_PACK_NT=$nt ZB_PACK_CHUNK_LOG2=$lg ZB_UPLOAD_RAW_EVERY=0 python tools/upload_bench.py 28
    run ZB_PACK_NT=$nt ZB_PACK_CHUNK_LOG2=$lg python tools/upload_bench.py 28
  done
done
