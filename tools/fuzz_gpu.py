#!/usr/bin/env python
"""Randomised differential test on the GPU: random tables / degrees / sizes / option settings / interleavings, every
result compared bit for bit with the oracle.  usage: python tools/fuzz_gpu.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zigz_b200 as z
from oracle import pyoracle as po

BB = z.BABYBEAR_P
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
L = z.lib()
t_end = time.time() + budget
stats = {"prove": 0, "fold_chain": 0, "eval": 0, "merkle": 0, "lasso": 0, "commitments": 0}


def rand_table(n):
    kind = rng.integers(0, 4)
    if kind == 0:
        return rng.integers(0, BB, size=n, dtype=np.uint64)
    if kind == 1:
        return np.full(n, BB - 1, dtype=np.uint64)
    if kind == 2:
        return (rng.integers(0, 2, size=n, dtype=np.uint64) * np.uint64(BB - 1))
    e = np.zeros(n, np.uint64)
    e[rng.integers(0, n)] = rng.integers(0, BB)
    return e


with z.Context(0) as ctx:
    while time.time() < t_end:
        ctx.set_option("tail_log2", int(rng.choice([0, 2, 5, 10, 14, 16])))
        ctx.set_option("prelaunch", int(rng.integers(0, 2)))
        L.zh_set_grid_min_log2(int(rng.choice([-1, 0, 3, 6, 10, 18])))
        ctx.set_option("prod_host_tail_log2", int(rng.choice([0, 1, 3, 7, 10, 12])))  # small product tables finish on the host
        ctx.set_option("host_tail_log2", int(rng.choice([2, 5, 10, 12])))            # d = 1: published-table size
        ctx.set_option("linear_d1", int(rng.integers(0, 2)))
        op = rng.integers(0, 6)
        if op == 0:  # full proofs
            d, lg = int(rng.integers(1, 4)), int(rng.integers(1, 19))
            es = [rand_table(1 << lg) for _ in range(d)]
            polys = [z.Multilinear.init(ctx, e) for e in es]
            want = po.prodcheck_prove(BB, es)
            for consume in (False, True):
                pr = z.ProductSumcheckProver.prove(polys, consume=consume)
                assert pr.round_polynomials.tolist() == want.round_polys.tolist(), ("prove", d, lg, consume)
                assert pr.final_evals == want.final_evals and pr.claimed_sum == want.claimed_sum
            for p in polys:
                p.deinit()
            stats["prove"] += 1
        elif op == 1:  # manual fold chains interleaved with other calls on two tables
            lg = int(rng.integers(2, 15))
            ea, eb = rand_table(1 << lg), rand_table(1 << lg)
            a, b = z.Multilinear.init(ctx, ea), z.Multilinear.init(ctx, eb)
            for _ in range(lg):
                which = rng.integers(0, 3)
                r = int(rng.integers(0, BB))
                if which == 0:
                    a.fold_inplace(r)
                    ea = po.mle_partial_eval(BB, ea, r)
                elif which == 1 and len(b) > 1:
                    b.fold_inplace(r)
                    eb = po.mle_partial_eval(BB, eb, r)
                else:
                    assert a.sum_over_hypercube() == po.mle_sum(BB, ea)
                if len(a) == 1:
                    break
            assert np.array_equal(a.evaluations, ea) and np.array_equal(b.evaluations, eb)
            a.deinit(); b.deinit()
            stats["fold_chain"] += 1
        elif op == 2:
            lg = int(rng.integers(0, 21))
            e = rand_table(1 << lg)
            p = z.Multilinear.init(ctx, e)
            pt = rng.integers(0, BB, size=lg, dtype=np.uint64)
            if lg <= 14:
                assert p.eval(pt) == po.mle_eval(BB, e, pt), ("eval", lg)
            else:  # oracle O(N v) too slow: linearity + boolean point
                idx = int(rng.integers(0, 1 << lg))
                assert p.eval([(idx >> k) & 1 for k in range(lg)]) == int(e[idx])
            p.deinit()
            stats["eval"] += 1
        elif op == 3:
            n = int(rng.integers(1, 1 << int(rng.integers(1, 14))))
            vals = rand_table(max(n, 1))[:n] if n > 0 else None
            t, w = z.SimpleMerkleTree.build(ctx, vals), po.merkle_build(vals)
            assert t.get_root() == w.root, ("merkle", n)
            idx = int(rng.integers(0, n))
            pr = t.open(idx)
            v, sib, dirs = po.merkle_open(w, idx)
            assert pr.value == v and np.array_equal(pr.path.siblings, sib)
            t.deinit()
            stats["merkle"] += 1
        elif op == 4:
            bits = int(rng.integers(1, 5))
            code = int(rng.integers(0, 3))
            table = po.build_table(BB, code, bits)
            nq = int(rng.integers(2, 600))
            q = table[rng.integers(0, table.shape[0], size=nq)]
            got, want = z.LassoProver.prove(ctx, table, q), po.lasso_prove(BB, table, q)
            assert got.sumcheck_proof.round_polynomials.tolist() == want.sumcheck.round_polys.tolist()
            assert (got.query_commitment, got.table_commitment) == (want.query_commitment, want.table_commitment)
            stats["lasso"] += 1
        else:
            lg, k = int(rng.integers(0, 9)), int(rng.integers(1, 70))
            es = [rand_table(1 << lg) for _ in range(k)]
            polys = [z.Multilinear.init(ctx, e) for e in es]
            tr, otr = z.FiatShamirTranscript(), po.Transcript()
            c, w = z.generate_commitments(tr, polys), po.generate_commitments(BB, otr, es)
            assert np.array_equal(c.roots, w.roots) and np.array_equal(c.values, w.values) and np.array_equal(c.siblings, w.siblings)
            assert tr.challenge() == otr.challenge(BB)
            for p in polys:
                p.deinit()
            stats["commitments"] += 1
    starved = ctx.get_option("starved")
print("fuzz ok", stats, "starved", starved, "seed", seed)
