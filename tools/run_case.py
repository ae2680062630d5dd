#!/usr/bin/env python
"""Runs ONE hot-path case a few times (for ncu / compute-sanitizer / quick timing on the GPU box).
usage: python tools/run_case.py {prod3|sumcheck|eval|merkle|lasso} [--log2n N] [--reps R]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import zigz_b200 as z  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("case")
ap.add_argument("--log2n", type=int, default=24)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--count", type=int, default=1, help="merkle: polynomials per batch")
ap.add_argument("--noprofile", action="store_true", help="no per-kernel events (latency measurements)")
a = ap.parse_args()
n = 1 << a.log2n
with z.Context(0) as ctx:
    if a.case == "prod3":
        polys = [z.Multilinear.synthetic(ctx, 1 + k, n) for k in range(3)]
        fn = lambda: z.ProductSumcheckProver.prove(polys)
    elif a.case == "sumcheck":
        poly = z.Multilinear.synthetic(ctx, 1, n)
        fn = lambda: z.SumcheckProver.prove(poly)
    elif a.case == "eval":
        poly = z.Multilinear.synthetic(ctx, 1, n)
        pt = np.arange(1, a.log2n + 1, dtype=np.uint64) * 7919 % z.BABYBEAR_P
        fn = lambda: poly.eval(pt)
    elif a.case == "merkle":
        polys = [z.Multilinear.synthetic(ctx, 1 + k, n) for k in range(a.count)]

        def fn():
            coms, trees = z.CommitmentScheme.batch_commit(polys)
            for t in trees:
                t.deinit()
    elif a.case == "lasso":
        rng = np.random.default_rng(1)
        x, y = rng.integers(0, 256, size=n, dtype=np.uint64), rng.integers(0, 256, size=n, dtype=np.uint64)
        q = np.ascontiguousarray(np.stack([x, y, x ^ y], axis=1))
        fn = lambda: z.LassoProver.prove_builtin(ctx, z.TABLE_XOR, 8, q)
    else:
        raise SystemExit("unknown case")
    fn()
    fn()
    ctx.profile(not a.noprofile)
    ctx.sync()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(a.reps):
        fn()
    ms = ctx.timer_stop()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"{a.case} 2^{a.log2n}: {ms / a.reps:.4f} ms/iter device, {wall / a.reps:.4f} ms/iter wall")
    for k, (cnt, kms, by) in sorted(ctx.profile_read().items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:16s} launches {cnt:5d}  total {kms:9.4f} ms  {by / (kms * 1e-3) / 1e9 if kms else 0:8.1f} GB/s (algorithmic)")
