#!/bin/bash
# One 8-GPU call: mask-context test on 8 GPUs, single-process strong scaling, torchrun strong/weak bench with gather sweeps
N=${1:-8}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mask_ctx.py -m gpu -x -q > gpurun_out/r02_gputest_n${N}.log 2>&1; tail -3 gpurun_out/r02_gputest_n${N}.log
for g in 2 4 8; do [ $g -le $N ] && python tools/mask_bench.py $g 30 5 >> gpurun_out/r02_mask_bench.jsonl 2>> gpurun_out/r02_mask_bench.err; done
python tools/mask_bench.py 1 30 5 >> gpurun_out/r02_mask_bench.jsonl 2>> gpurun_out/r02_mask_bench.err
cat gpurun_out/r02_mask_bench.jsonl
for gl in 16 13 10; do
  ZB_GATHER_LOG2=$gl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((gl % 10)) bench.py --gpus $N --steps 10 --warmup 3 --skip-e2e > gpurun_out/r02_bench_n${N}_gather$gl.json 2> gpurun_out/r02_bench_n${N}_gather$gl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n${N}_gather$gl.json"))
    s=d["extras"]["C5_strong_2^30_total"]
    print("gather $gl: weak", round(d["ms_per_step"],3), "ms; strong", round(s["ms_per_step"],3), "ms", s["exchange"])
except Exception as e:
    print("gather $gl failed", e)
PY
done
