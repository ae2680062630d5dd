"""CUDA Multilinear kernels (through the C ABI) against the oracle. Bit-exact: integer arithmetic mod p."""
import numpy as np
import pytest

from _cases import BB, synthetic

pytestmark = pytest.mark.gpu


def rand_evals(rng, n):
    e = rng.integers(0, BB, size=n, dtype=np.uint64)
    if n >= 4:  # extremes
        e[0], e[-1], e[n // 2] = BB - 1, BB - 1, 0
    return e


def test_reference_unit_test_facts(zlib, ctx):  # src/poly/multilinear.zig:337-566 (values < 17 behave alike in BabyBear)
    p = zlib.Multilinear.init(ctx, [1, 2, 3, 4])
    assert p.num_vars == 2 and len(p) == 4
    assert p.evaluations.tolist() == [1, 2, 3, 4]
    assert [p.eval(pt) for pt in ([0, 0], [1, 0], [0, 1], [1, 1])] == [1, 2, 3, 4]  # :383-413
    q = p.partial_eval(0)  # :436-463
    assert q.num_vars == 1 and q.evaluations.tolist() == [1, 2]
    assert p.sum_over_hypercube() == 10  # :465-477
    assert p.round_polynomial() == [3, 4]  # :479-506
    x = zlib.Multilinear.init(ctx, [0, 1])  # :415-434  p(x) = x
    assert x.eval([2]) == 2 and x.eval([5]) == 5
    z = zlib.Multilinear.zero(ctx, 3)  # :359-368
    assert z.evaluations.tolist() == [0] * 8
    c = zlib.Multilinear.constant(ctx, 2, 5)  # :370-381
    assert c.evaluations.tolist() == [5] * 4 and c.eval([123, 456]) == 5
    a = zlib.Multilinear.init(ctx, [5, 6, 7, 8])
    assert p.add(a).evaluations.tolist() == [6, 8, 10, 12]  # :508-527
    assert p.scalar_mul(2).evaluations.tolist() == [2, 4, 6, 8]  # :529-544


def test_init_errors(zlib, ctx):  # multilinear.zig:36-47
    with pytest.raises(zlib.ZigzError) as e:
        zlib.Multilinear.init(ctx, [1, 2, 3])
    assert e.value.name == "LengthNotPowerOfTwo"
    with pytest.raises(zlib.ZigzError) as e:
        zlib.Multilinear.init(ctx, np.zeros(0, np.uint64))
    assert e.value.name == "EmptyEvaluations"
    with pytest.raises(zlib.ZigzError) as e:  # the reference asserts val < MODULUS (field.zig:43)
        zlib.Multilinear.init(ctx, [1, BB, 3, 4])
    assert e.value.name == "NotCanonical"
    p = zlib.Multilinear.init(ctx, [1, 2, 3, 4])
    with pytest.raises(zlib.ZigzError) as e:
        p.eval([1])
    assert e.value.name == "WrongNumberOfVariables"  # :112
    one = zlib.Multilinear.init(ctx, [7])
    assert one.num_vars == 0 and one.eval([]) == 7
    for fn in (lambda: one.partial_eval(3), one.round_polynomial):
        with pytest.raises(zlib.ZigzError) as e:
            fn()
        assert e.value.name == "NoVariables"  # :156, :207
    with pytest.raises(zlib.ZigzError) as e:
        p.add(one)
    assert e.value.name == "DifferentNumberOfVariables"  # :237


@pytest.mark.parametrize("lg", list(range(0, 15)) + [17, 20])
def test_sum_round_poly_partial_eval_vs_oracle(zlib, ctx, po, lg):
    rng = np.random.default_rng(lg)
    n = 1 << lg
    e = rand_evals(rng, n)
    p = zlib.Multilinear.init(ctx, e)
    assert np.array_equal(p.evaluations, e)
    assert p.sum_over_hypercube() == po.mle_sum(BB, e)
    if lg == 0:
        return
    assert p.round_polynomial() == po.mle_round_poly(BB, e)
    for r in (0, 1, BB - 1, int(rng.integers(0, BB))):
        q, nxt = p.partial_eval(r, with_next_sums=True)
        want = po.mle_partial_eval(BB, e, r)
        assert np.array_equal(q.evaluations, want), (lg, r)
        if lg >= 2:
            rp = po.mle_round_poly(BB, want)
            assert [nxt[0], (nxt[1] - nxt[0]) % BB] == rp
        else:
            assert nxt[0] == want[0]
        q.deinit()
    assert np.array_equal(p.evaluations, e)  # partialEval leaves its input untouched


@pytest.mark.parametrize("lg", [1, 2, 3, 4, 5, 9, 13, 16])
def test_fold_inplace_chain_vs_oracle(zlib, ctx, po, lg):
    rng = np.random.default_rng(100 + lg)
    e = rand_evals(rng, 1 << lg)
    p = zlib.Multilinear.init(ctx, e)
    cur = e
    for _ in range(lg):
        r = int(rng.integers(0, BB))
        nxt = p.fold_inplace(r)
        cur = po.mle_partial_eval(BB, cur, r)
        assert len(p) == cur.size
        assert np.array_equal(p.evaluations, cur)
        if cur.size >= 2:
            rp = po.mle_round_poly(BB, cur)
            assert [nxt[0], (nxt[1] - nxt[0]) % BB] == rp
        else:
            assert nxt[0] == cur[0]


@pytest.mark.parametrize("lg", list(range(0, 17)))
def test_eval_vs_oracle(zlib, ctx, po, lg):
    """O(N) LSB-first fold on the device == the reference's O(N v) basis-product loop."""
    rng = np.random.default_rng(200 + lg)
    e = rand_evals(rng, 1 << lg)
    p = zlib.Multilinear.init(ctx, e)
    for _ in range(3):
        pt = rng.integers(0, BB, size=lg, dtype=np.uint64)
        assert p.eval(pt) == po.mle_eval(BB, e, pt), lg
    if lg:
        idx = int(rng.integers(0, 1 << lg))  # boolean point -> stored value
        assert p.eval([(idx >> k) & 1 for k in range(lg)]) == int(e[idx])
        assert p.eval([BB - 1] * lg) == po.mle_eval(BB, e, [BB - 1] * lg)


def test_eval_golden(zlib, ctx, golden):
    for name, case in golden["eval"].items():
        if "seed" in case:
            p = zlib.Multilinear.init(ctx, synthetic(case["seed"], case["n"]))
            assert p.eval(case["point"]) == case["value"], name


def test_eval_large_linearity(zlib, ctx):
    """Full-size property (2^24): eval is linear in the evaluations and reproduces stored values at boolean points."""
    lg = 24
    a = zlib.Multilinear.synthetic(ctx, 1, 1 << lg)
    b = zlib.Multilinear.synthetic(ctx, 2, 1 << lg)
    s = a.add(b)
    rng = np.random.default_rng(7)
    pt = rng.integers(0, BB, size=lg, dtype=np.uint64)
    assert s.eval(pt) == (a.eval(pt) + b.eval(pt)) % BB
    k = 123456789
    assert a.scalar_mul(k).eval(pt) == a.eval(pt) * k % BB
    idx = 0xABCDE5
    from _cases import splitmix64
    assert a.eval([(idx >> j) & 1 for j in range(lg)]) == splitmix64(1 + idx) % BB
    for m in (a, b, s):
        m.deinit()


def test_synthetic_matches_oracle_generator(zlib, ctx, po):
    m = zlib.Multilinear.synthetic(ctx, 0x5A49475A, 1 << 12)
    assert np.array_equal(m.evaluations, po.fill_synthetic(BB, 0x5A49475A, 0, 1 << 12))
    # cyclic shard g of P: element j of the shard is global index g + P*j
    full = po.fill_synthetic(BB, 9, 0, 1 << 10)
    for g in range(4):
        sh = zlib.Multilinear.synthetic(ctx, 9, 1 << 8, start=g, stride=4)
        assert np.array_equal(sh.evaluations, full[g::4])


def test_u32_upload_and_clone(zlib, ctx):
    e = synthetic(5, 1 << 10)
    a = zlib.Multilinear.init_u32(ctx, e.astype(np.uint32))
    b = a.clone()
    a.fold_inplace(12345)
    assert np.array_equal(b.evaluations, e)
    with pytest.raises(zlib.ZigzError) as err:
        zlib.Multilinear.init_u32(ctx, np.array([1, 2, 3, BB + 5], np.uint32))
    assert err.value.name == "NotCanonical"


@pytest.mark.parametrize("lg,count", [(0, 3), (1, 2), (3, 5), (4, 1), (9, 43), (12, 4), (13, 3), (17, 2), (20, 3), (21, 2), (23, 2)])
def test_eval_batch_matches_single_evaluations(zlib, ctx, po, lg, count):
    """zb_mle_eval_batch (all opening points of Prover.generateCommitments at once, prover.zig:420-443) == count x eval."""
    import ctypes as C
    polys = [zlib.Multilinear.synthetic(ctx, 3000 + 17 * lg + i, 1 << lg) for i in range(count)]
    pts = np.ascontiguousarray(po.fill_synthetic(BB, 77 + lg, 0, max(count * lg, 1))[: count * lg].reshape(count, lg))
    hs = (C.c_uint64 * count)(*[p.handle for p in polys])
    out = np.zeros(count, np.uint64)
    ctx.check(zlib.lib().zb_mle_eval_batch(ctx.handle, hs, count, pts.ctypes.data_as(zlib.api.P64) if lg else None, lg,
                                           out.ctypes.data_as(zlib.api.P64)))
    for i, p in enumerate(polys):
        assert int(out[i]) == p.eval(pts[i]), (lg, i)
        if lg <= 12:
            assert int(out[i]) == po.mle_eval(BB, p.evaluations, pts[i])
        p.deinit()


def test_merkle_open_batch_matches_single_openings(zlib, ctx, po):
    import ctypes as C
    for lg, count in ((0, 2), (1, 3), (6, 43), (11, 5)):
        polys = [zlib.Multilinear.synthetic(ctx, 5000 + i, 1 << lg) for i in range(count)]
        coms, trees = zlib.CommitmentScheme.batch_commit(polys)
        idx = np.array([(i * 2654435761) % (1 << lg) for i in range(count)], np.uint64)
        hs = (C.c_uint64 * count)(*[t.handle for t in trees])
        h = max(lg, 1)
        sib, dirs, vals = np.zeros((count, h, 32), np.uint8), np.zeros((count, h), np.uint8), np.zeros(count, np.uint64)
        ctx.check(zlib.lib().zb_merkle_open_batch(ctx.handle, hs, count, idx.ctypes.data_as(zlib.api.P64), sib.ctypes.data_as(zlib.api.P8),
                                                  dirs.ctypes.data_as(zlib.api.P8), vals.ctypes.data_as(zlib.api.P64)))
        sib, dirs = sib.reshape(-1)[: count * lg * 32].reshape(count, lg, 32), dirs.reshape(-1)[: count * lg].reshape(count, lg)
        for i, t in enumerate(trees):
            pr = t.open(int(idx[i]))
            assert pr.value == int(vals[i])
            assert np.array_equal(pr.path.siblings, sib[i]) and np.array_equal(pr.path.directions, dirs[i])
        bad = idx.copy()
        bad[-1] = 1 << lg
        assert zlib.lib().zb_merkle_open_batch(ctx.handle, hs, count, bad.ctypes.data_as(zlib.api.P64), None, None, None) == -6  # IndexOutOfBounds
        for t in trees:
            t.deinit()
        for p in polys:
            p.deinit()
