#!/usr/bin/env python
"""Single-launch bandwidth of the d=1 kernels at a large size: roundPolynomial, partialEval (+ next sums), sum."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zigz_b200 as z
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 29
with z.Context(0) as ctx:
    n = 1 << lg
    p = z.Multilinear.synthetic(ctx, 1, n)
    def t(fn, reps=10):
        fn(); ctx.sync(); ctx.timer_start()
        for _ in range(reps): fn()
        return ctx.timer_stop() / reps
    def pe():
        q = p.partial_eval(12345); q.deinit()
    ms = t(pe); print(f"partial_eval d=1 2^{lg}: {ms:.4f} ms  {6*n/ms/1e6:.0f} GB/s")
    ms = t(p.round_polynomial); print(f"round_polynomial d=1 2^{lg}: {ms:.4f} ms  {4*n/ms/1e6:.0f} GB/s")
    ms = t(p.sum_over_hypercube); print(f"sum 2^{lg}: {ms:.4f} ms  {4*n/ms/1e6:.0f} GB/s")
