#!/usr/bin/env python
"""Where a Lasso proof's time goes: row upload + XXH3, sumcheck, the two flat SHA3 commitments (host sponge)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zigz_b200 as z
from zigz_b200.api import _p64, _p8

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
with z.Context(0) as ctx:
    L = z.lib()
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=1 << lg, dtype=np.uint64)
    b = rng.integers(0, 256, size=1 << lg, dtype=np.uint64)
    q = np.ascontiguousarray(np.stack([a, b, a ^ b], axis=1))
    for rep in range(3):
        t = [time.perf_counter()]
        tab = C.c_uint64(0)
        ctx.check(L.zb_table_mle(ctx.handle, z.TABLE_XOR, 8, C.byref(tab)))
        ctx.sync(); t.append(time.perf_counter())
        qp = C.c_uint64(0)
        ctx.check(L.zb_xxh3_rows(ctx.handle, _p64(q.reshape(-1)), q.shape[0], 3, q.shape[0], C.byref(qp)))
        ctx.sync(); t.append(time.perf_counter())
        poly = z.Multilinear(ctx, qp.value)
        proof = z.SumcheckProver.prove(poly)
        t.append(time.perf_counter())
        qc, tc = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
        ctx.check(L.zh_lasso_commit_poly(ctx.handle, qp.value, _p8(qc)))
        t.append(time.perf_counter())
        ctx.check(L.zh_lasso_commit_poly(ctx.handle, tab.value, _p8(tc)))
        t.append(time.perf_counter())
        poly.deinit()
        L.zb_mle_free(ctx.handle, tab.value)
        t0 = time.perf_counter()
        z.LassoProver.prove_builtin(ctx, z.TABLE_XOR, 8, q)
        whole = time.perf_counter() - t0
        d = [(t[i + 1] - t[i]) * 1e3 for i in range(len(t) - 1)]
        print(f"2^{lg} lookups: table_mle {d[0]:.2f} ms | upload+xxh3 {d[1]:.2f} | sumcheck {d[2]:.2f} | query commit {d[3]:.2f} "
              f"({d[3] * 1e6 / ((1 << lg) / 17):.0f} ns/perm) | table commit {d[4]:.2f} | whole prove {whole * 1e3:.2f}")
