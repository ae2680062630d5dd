#!/usr/bin/env python
"""One process, several GPUs (zb_ctx_create_mask): time the degree-3 product sumcheck of ONE 2^LOG2N-entry job sharded over the
GPUs of a device mask, and the sharded Merkle commit of one table.  usage: python tools/mask_bench.py <n_gpus> [log2n] [reps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zigz_b200 as z  # noqa: E402

gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
SEED = 0x5A49475A
out = {"n_gpus": gpus, "log2_n_total": lg}
with z.Context(device_mask=(1 << gpus) - 1) as ctx:
    polys = [z.Multilinear.synthetic(ctx, SEED + k, 1 << lg) for k in range(3)]
    for _ in range(2):
        pr = z.ProductSumcheckProver.prove(polys)
    t0 = time.perf_counter()
    for _ in range(reps):
        pr = z.ProductSumcheckProver.prove(polys)
    dt = (time.perf_counter() - t0) / reps
    out["prodcheck_d3"] = {"ms_per_prove": dt * 1e3, "melem_per_s": (1 << lg) / dt / 1e6, "claimed_sum": pr.claimed_sum,
                           "final_evals": list(pr.final_evals)}
    if "--no-merkle" in sys.argv:
        print(json.dumps(out))
        sys.exit(0)
    lgm = min(lg, 26 + (gpus - 1).bit_length())
    pm = polys[0] if lgm == lg else z.Multilinear.synthetic(ctx, SEED + 9, 1 << lgm)
    com, tree = z.CommitmentScheme.commit(pm)
    tree.deinit()
    t0 = time.perf_counter()
    com, tree = z.CommitmentScheme.commit(pm)
    dm = time.perf_counter() - t0
    t0 = time.perf_counter()
    op = tree.open(12345)
    do = time.perf_counter() - t0
    out["merkle_commit"] = {"log2_leaves": lgm, "ms": dm * 1e3, "keccak_per_s": (2 * (1 << lgm) - 1) / dm, "open_ms": do * 1e3,
                            "root": com.commitment.hex()}
print(json.dumps(out))
