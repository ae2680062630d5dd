#!/usr/bin/env python
"""Every kernel of the library once, at small sizes, checked against the oracle — for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zigz_b200 as z
from oracle import pyoracle as po
BB = z.BABYBEAR_P
with z.Context(0) as ctx:
    for tail in (14, 0):
        ctx.set_option("tail_log2", tail)
        for d, lg in ((1, 11), (2, 9), (3, 10), (3, 3), (1, 1)):
            es = [po.fill_synthetic(BB, 5 + k, 0, 1 << lg) for k in range(d)]
            polys = [z.Multilinear.init(ctx, e) for e in es]
            want = po.prodcheck_prove(BB, es)
            for consume in (False, True):
                pr = z.ProductSumcheckProver.prove(polys, consume=consume)
                assert pr.round_polynomials.tolist() == want.round_polys.tolist() and pr.final_evals == want.final_evals
    e = po.fill_synthetic(BB, 9, 0, 1 << 20)   # eval: warp kernel + block kernel
    p = z.Multilinear.init_u32(ctx, e.astype(np.uint32))
    pt = po.fill_synthetic(BB, 10, 0, 20)
    ev = p.eval(pt)
    small = z.Multilinear.init(ctx, e[:1 << 9])
    assert small.eval(pt[:9]) == po.mle_eval(BB, e[:1 << 9], pt[:9])
    assert p.sum_over_hypercube() == int(e.sum() % BB)
    a, b = z.Multilinear.init(ctx, e[:64]), z.Multilinear.init(ctx, e[64:128])
    assert a.add(b).evaluations.tolist() == ((e[:64] + e[64:128]) % BB).tolist()
    assert a.scalar_mul(7).evaluations.tolist() == (e[:64] * 7 % BB).tolist()
    for n in (1, 5, 1000, 4096):
        vals = e[:n]
        t = z.SimpleMerkleTree.build(ctx, vals)
        w = po.merkle_build(vals)
        assert t.get_root() == w.root
        pr = t.open(n - 1)
        assert z.SimpleMerkleTree.verify(t.get_root(), pr)
    polys = [z.Multilinear.init(ctx, po.fill_synthetic(BB, 30 + i, 0, 256)) for i in range(43)]
    tr, otr = z.FiatShamirTranscript(), po.Transcript()
    c = z.generate_commitments(tr, polys)
    w = po.generate_commitments(BB, otr, [q.evaluations for q in polys])
    assert np.array_equal(c.roots, w.roots) and np.array_equal(c.values, w.values) and np.array_equal(c.siblings, w.siblings)
    table = po.build_table(BB, po.TABLE_ADD, 4)
    q = table[np.arange(300) % 256]
    lp, wl = z.LassoProver.prove(ctx, table, q), po.lasso_prove(BB, table, q)
    assert (lp.query_commitment, lp.table_commitment) == (wl.query_commitment, wl.table_commitment)
    lb = z.LassoProver.prove_builtin(ctx, z.TABLE_ADD, 4, q)
    assert lb.table_commitment == wl.table_commitment
    cols = np.arange(43 * 37, dtype=np.uint64).reshape(43, 37) * 2654435761
    wp = z.witness_pack(ctx, cols)
    assert np.array_equal(np.stack([m.evaluations for m in wp]), po.witness_pack(BB, cols, 33))
    print("sanity_small ok;", ctx.kernel_launches, "launches; eval", ev)
