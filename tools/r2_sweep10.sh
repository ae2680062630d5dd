#!/bin/bash
# where to leave the sharded regime, and how (in-kernel gather over the exchange buffers vs host rendezvous + direct peer reads)
# usage: bash tools/r2_sweep10.sh <n_gpus> <tag>
N=$1
out=gpurun_out/r02_sweep10_$2.txt
: > $out
run() { echo "## $*" >> $out; env "$@" 2>/dev/null | cut -c1-140 >> $out; }
run ZB_GATHER_XCHG=1 ZB_GATHER_LOG2=16 python tools/mask_bench.py $N 30 20 --no-merkle
run ZB_GATHER_XCHG=1 ZB_GATHER_LOG2=13 python tools/mask_bench.py $N 30 20 --no-merkle
run ZB_GATHER_XCHG=1 ZB_GATHER_LOG2=11 python tools/mask_bench.py $N 30 20 --no-merkle
run ZB_GATHER_XCHG=0 ZB_GATHER_LOG2=16 python tools/mask_bench.py $N 30 20 --no-merkle
