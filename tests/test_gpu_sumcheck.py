"""SumcheckProver on the GPU (device rounds + host transcript) against the golden vectors and the oracle."""
import hashlib

import numpy as np
import pytest

from _cases import BB, assert_sumcheck_equal, sumcheck_case_evals, synthetic

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("rounds_mode")]


def test_golden_sumcheck(zlib, ctx, golden):
    for name, case in golden["sumcheck"].items():
        if case["p"] != BB:
            continue
        poly = zlib.Multilinear.init(ctx, sumcheck_case_evals(case))
        if "challenges" in case:
            pr = zlib.SumcheckProver.prove_interactive(poly, case["challenges"])
        else:
            pr = zlib.SumcheckProver.prove(poly)
            assert pr.claimed_sum == case["claimed_sum"], name
        assert_sumcheck_equal(pr, case)
        assert hashlib.sha3_256(pr.to_bytes()).hexdigest() == case["to_bytes_sha3"], name


def test_f17_reference_example_shape(zlib, ctx, po):
    # examples/sumcheck_basic.zig uses F17 [1,2,3,4]; in BabyBear the first round polynomial is the same [3, 4]
    pr = zlib.SumcheckProver.prove(zlib.Multilinear.init(ctx, [1, 2, 3, 4]))
    want = po.sumcheck_prove(BB, [1, 2, 3, 4])
    assert pr.round_polynomials[0].tolist() == [3, 4] and pr.claimed_sum == 10
    assert pr.to_bytes() == want.to_bytes()


@pytest.mark.parametrize("lg", list(range(1, 14)) + [16, 18, 20])
def test_prove_vs_oracle(zlib, ctx, po, lg):
    e = po.fill_synthetic(BB, 0x5A49475A, 0, 1 << lg)
    poly = zlib.Multilinear.synthetic(ctx, 0x5A49475A, 1 << lg)
    pr = zlib.SumcheckProver.prove(poly)
    want = po.sumcheck_prove(BB, e)
    assert pr.claimed_sum == want.claimed_sum
    assert pr.round_polynomials.tolist() == want.round_polys.tolist()
    assert pr.final_point.tolist() == want.final_point.tolist()
    assert pr.final_eval == want.final_eval
    assert pr.to_bytes() == want.to_bytes()
    assert np.array_equal(poly.evaluations[:64], e[:64])  # prove() leaves the polynomial untouched
    poly.deinit()


def test_scalability_example_pattern(zlib, ctx, po):  # examples/sumcheck_scalability.zig:44-46: e[i] = i + 1
    for lg in range(1, 9):
        e = np.arange(1, (1 << lg) + 1, dtype=np.uint64)
        pr = zlib.SumcheckProver.prove(zlib.Multilinear.init(ctx, e))
        assert pr.to_bytes() == po.sumcheck_prove(BB, e).to_bytes()


def test_prove_interactive_and_errors(zlib, ctx, po):
    e = synthetic(3, 1 << 9)
    ch = synthetic(4, 9)
    poly = zlib.Multilinear.init(ctx, e)
    pr = zlib.SumcheckProver.prove_interactive(poly, ch)
    want = po.sumcheck_prove_interactive(BB, e, ch)
    assert pr.round_polynomials.tolist() == want.round_polys.tolist() and pr.final_eval == want.final_eval
    assert pr.final_point.tolist() == ch.tolist()
    with pytest.raises(zlib.ZigzError) as err:
        zlib.SumcheckProver.prove_interactive(poly, ch[:5])
    assert err.value.name == "WrongNumberOfChallenges"  # sumcheck_prover.zig:105
    with pytest.raises(zlib.ZigzError) as err:
        zlib.SumcheckProver.prove(zlib.Multilinear.init(ctx, [5]))
    assert err.value.name == "NoVariables"  # sumcheck_prover.zig:31


def test_golden_prodcheck(zlib, ctx, golden):
    for name, case in golden["prodcheck"].items():
        polys = [zlib.Multilinear.init(ctx, synthetic(case["seed"] + k, case["n"])) for k in range(case["d"])]
        pr = zlib.ProductSumcheckProver.prove(polys)
        assert pr.claimed_sum == case["claimed_sum"], name
        assert_sumcheck_equal(pr, case)


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("lg", [1, 2, 3, 4, 5, 7, 10, 12, 15])
def test_prodcheck_vs_oracle(zlib, ctx, po, d, lg):
    n = 1 << lg
    es = [po.fill_synthetic(BB, 0x5A49475A + k, 0, n) for k in range(d)]
    polys = [zlib.Multilinear.synthetic(ctx, 0x5A49475A + k, n) for k in range(d)]
    want = po.prodcheck_prove(BB, es)
    for consume in (False, True):
        pr = zlib.ProductSumcheckProver.prove(polys, consume=consume)
        assert pr.claimed_sum == want.claimed_sum, (d, lg)
        assert pr.round_polynomials.tolist() == want.round_polys.tolist(), (d, lg)
        assert pr.final_point.tolist() == want.final_point.tolist()
        assert pr.final_evals == want.final_evals
    assert all(len(p) == 1 for p in polys)  # consumed
    if d == 1:  # degree 1 reduces to the reference prover exactly
        ref = po.sumcheck_prove(BB, es[0])
        assert pr.round_polynomials.tolist() == ref.round_polys.tolist() and pr.final_evals[0] == ref.final_eval


@pytest.mark.parametrize("d,lg", [(1, 24), (3, 22)])
def test_full_size_properties(zlib, ctx, po, d, lg):
    """At sizes the oracle cannot follow in seconds: the proof must pass the reference verifier's round checks
    (g_r(0) + g_r(1) == claim, claim <- g_r(challenge)) with the verifier re-deriving the same challenges, and the
    final claim must equal the product of the final evaluations (verifyRounds, sumcheck_verifier.zig:172-202)."""
    polys = [zlib.Multilinear.synthetic(ctx, 77 + k, 1 << lg) for k in range(d)]
    pr = zlib.ProductSumcheckProver.prove(polys, consume=True)
    ok, final_claim = po.sumcheck_verify_rounds(BB, pr.round_polynomials, pr.claimed_sum)
    assert ok
    prod = 1
    for x in pr.final_evals:
        prod = prod * x % BB
    assert final_claim == prod
    # the verifier's transcript replay gives the prover's challenges
    t = po.Transcript()
    for r in range(lg):
        for c in pr.round_polynomials[r]:
            t.append_field(int(c))
        assert t.challenge(BB) == int(pr.final_point[r])
    for p in polys:
        p.deinit()


@pytest.mark.parametrize("d,lg", [(3, 30), (1, 28), (1, 32), (2, 31)])
def test_baseline_size_fold_chain_agrees_with_eval_kernel(zlib, ctx, po, d, lg):
    """BASELINE config C5 at its full size (three 2^30-entry tables), the d=1 prover at 2^28, and tables of 2^31 / 2^32
    entries (element indices and byte offsets beyond 32 bits — the largest single tables 180 GB hold twice): besides the verifier's
    round checks, every final evaluation must equal Multilinear.eval of the untouched table at the challenge point —
    computed by a different kernel family (k_eval_*), LSB-first (multilinear.zig:110-144) where the prover binds the
    top index bit first (:154-180), hence the reversed point."""
    info = ctx.device_info()  # free memory counts the context's own cached blocks as used: compare against the total
    if info["total_mem"] < (2 * d + 2) * (4 << lg):
        pytest.skip("not enough free device memory for this size")
    polys = [zlib.Multilinear.synthetic(ctx, 500 + k, 1 << lg) for k in range(d)]
    pr = zlib.ProductSumcheckProver.prove(polys, consume=False)
    ok, final_claim = po.sumcheck_verify_rounds(BB, pr.round_polynomials, pr.claimed_sum)
    assert ok
    prod = 1
    for x in pr.final_evals:
        prod = prod * x % BB
    assert final_claim == prod
    point = [int(x) for x in pr.final_point][::-1]
    for k, p in enumerate(polys):
        assert p.eval(point) == pr.final_evals[k], k
        p.deinit()


# ---------------------------------------------------------------- persistent tail kernel (zb_set_option "tail_log2")
@pytest.mark.parametrize("prelaunch", [1, 0])
@pytest.mark.parametrize("tail_log2", [0, 2, 3, 10, 14, 20])
def test_tail_kernel_settings_give_identical_proofs(zlib, ctx, po, tail_log2, prelaunch):
    """Persistent tail kernel on/off at several thresholds x pre-launched fold kernels on/off: same proofs."""
    old, oldp = ctx.get_option("tail_log2"), ctx.get_option("prelaunch")
    try:
        ctx.set_option("tail_log2", tail_log2)
        ctx.set_option("prelaunch", prelaunch)
        ctx.set_option("linear_d1", prelaunch)  # d = 1: several-rounds-per-pass prover on / off as well
        for d, lg in ((1, 1), (1, 2), (1, 9), (1, 16), (3, 2), (3, 11), (2, 13), (3, 17)):
            es = [po.fill_synthetic(BB, 900 + k, 0, 1 << lg) for k in range(d)]
            polys = [zlib.Multilinear.init(ctx, e) for e in es]
            want = po.prodcheck_prove(BB, es)
            for consume in (False, True):
                pr = zlib.ProductSumcheckProver.prove(polys, consume=consume)
                assert pr.round_polynomials.tolist() == want.round_polys.tolist(), (tail_log2, d, lg)
                assert pr.final_evals == want.final_evals and pr.final_point.tolist() == want.final_point.tolist()
    finally:
        ctx.set_option("tail_log2", old)
        ctx.set_option("prelaunch", oldp)
        ctx.set_option("linear_d1", 1)
    assert ctx.get_option("starved") == 0  # no polling kernel ever left without its challenge


def test_tail_session_survives_interleaved_calls(zlib, ctx, po):
    """Any other call on the context ends the running persistent kernel; the tables must stay consistent with the
    rounds completed and folding must be resumable (new session)."""
    rng = np.random.default_rng(5)
    e = po.fill_synthetic(BB, 31, 0, 1 << 10)
    f = po.fill_synthetic(BB, 32, 0, 1 << 8)
    a, b = zlib.Multilinear.init(ctx, e), zlib.Multilinear.init(ctx, f)
    cur_a, cur_b = e, f
    for step in range(8):
        r = int(rng.integers(0, BB))
        nxt = a.fold_inplace(r)  # tail session on `a`
        cur_a = po.mle_partial_eval(BB, cur_a, r)
        rp = po.mle_round_poly(BB, cur_a)
        assert [nxt[0], (nxt[1] - nxt[0]) % BB] == rp
        if step % 3 == 0:
            assert np.array_equal(a.evaluations, cur_a)  # download: quiesces the session
        if step % 3 == 1:
            r2 = int(rng.integers(0, BB))
            b.fold_inplace(r2)  # a different table: session switches
            cur_b = po.mle_partial_eval(BB, cur_b, r2)
        if step % 3 == 2:
            assert a.round_polynomial() == rp
    assert np.array_equal(a.evaluations, cur_a) and np.array_equal(b.evaluations, cur_b)
    a.deinit()  # freeing a table with a live session
    assert np.array_equal(b.evaluations, cur_b)
    with pytest.raises(zlib.ZigzError):
        b.fold_inplace(BB)  # not canonical: rejected without disturbing the table
    assert np.array_equal(b.evaluations, cur_b)


def test_headline_size_properties(zlib, ctx, po):
    """BASELINE config C5 on one GPU (three 2^30-entry tables when memory allows, else 2^28): verifier round checks,
    transcript replay, final claim = product of final evaluations, and claimed sum against an independent device sum
    of the element-wise product structure (sum over the hypercube of g == g_0(0) + g_0(1))."""
    lg = 30 if ctx.device_info()["free_mem"] > 40 << 30 else 28
    polys = [zlib.Multilinear.synthetic(ctx, 0x5A49475A + k, 1 << lg) for k in range(3)]
    pr = zlib.ProductSumcheckProver.prove(polys)  # non-consuming: inputs stay intact
    assert pr.num_vars == lg
    ok, final_claim = po.sumcheck_verify_rounds(BB, pr.round_polynomials, pr.claimed_sum)
    assert ok
    prod = 1
    for x in pr.final_evals:
        prod = prod * x % BB
    assert final_claim == prod
    # the inputs were not modified, and a second prove is bit-identical (determinism)
    from _cases import splitmix64
    assert int(polys[1].evaluations[:4][3]) == splitmix64(0x5A49475A + 1 + 3) % BB
    pr2 = zlib.ProductSumcheckProver.prove(polys, consume=True)
    assert pr2.round_polynomials.tolist() == pr.round_polynomials.tolist() and pr2.final_evals == pr.final_evals
    for p in polys:
        p.deinit()


def test_tail_kernel_starvation_falls_back_to_launches(zlib, po):
    """If the persistent kernel gives up waiting for its first challenge (a profiler serialising launches, a stalled
    host thread) the round is redone with a plain launch, the context stops using the tail kernel, results unchanged."""
    with zlib.Context(0) as c2:
        c2.set_option("tail_log2", 14)  # the default, unless ZB_TAIL_LOG2 overrides it
        c2.set_option("prelaunch", 1)
        c2.set_option("linear_d1", 0)  # the one-round-per-kernel prover is the one with polling kernels
        c2.set_option("prod_host_tail_log2", 0)  # ... and it only runs when small tables do not finish on the host
        for lg in (12, 18):  # 2^12: the tail kernel starves; 2^18: a pre-launched fold kernel starves
            e = po.fill_synthetic(BB, 4242, 0, 1 << lg)
            poly = zlib.Multilinear.init(c2, e)
            want = po.sumcheck_prove(BB, e)
            before = c2.get_option("starved")
            c2.set_option("tail_test_starve", 1)
            assert zlib.SumcheckProver.prove(poly).to_bytes() == want.to_bytes()
            assert c2.get_option("starved") == before + 1
            assert c2.get_option("tail_log2") == 14  # a single hiccup does not switch the mechanism off
            assert zlib.SumcheckProver.prove(poly).to_bytes() == want.to_bytes()
        c2.set_option("starved", 2)
        c2.set_option("tail_test_starve", 1)
        assert zlib.SumcheckProver.prove(poly).to_bytes() == want.to_bytes()
        assert c2.get_option("tail_log2") == 0 and c2.get_option("prelaunch") == 0  # third time: switched off
        assert zlib.SumcheckProver.prove(poly).to_bytes() == want.to_bytes()


# ---------------------------------------------------------------- two rounds per pass (zb_prod_grid / zb_prod_fold_grid)
@pytest.mark.parametrize("host_tail", [0, 9])
@pytest.mark.parametrize("grid_min", [3, 5, 8, 0, -1])
def test_two_rounds_per_pass_gives_identical_proofs(zlib, ctx, po, grid_min, host_tail):
    """Force the bivariate-grid path down to tiny tables (and switch it off), with and without the host finishing the small
    tables: proofs must not change by a bit."""
    old = zlib.lib().zh_set_grid_min_log2(grid_min)
    old_ht = ctx.get_option("prod_host_tail_log2")
    ctx.set_option("prod_host_tail_log2", host_tail)
    try:
        for d, lg in ((1, 5), (1, 6), (1, 9), (1, 14), (2, 5), (2, 8), (2, 11), (3, 5), (3, 6), (3, 7), (3, 10), (3, 13), (3, 16)):
            es = [po.fill_synthetic(BB, 1234 + k, 0, 1 << lg) for k in range(d)]
            polys = [zlib.Multilinear.init(ctx, e) for e in es]
            want = po.prodcheck_prove(BB, es)
            for consume in (False, True):
                pr = zlib.ProductSumcheckProver.prove(polys, consume=consume)
                assert pr.claimed_sum == want.claimed_sum, (grid_min, d, lg)
                assert pr.round_polynomials.tolist() == want.round_polys.tolist(), (grid_min, d, lg, consume)
                assert pr.final_point.tolist() == want.final_point.tolist() and pr.final_evals == want.final_evals
                if not consume:
                    assert np.array_equal(polys[0].evaluations, es[0])  # non-consuming prove keeps the inputs
        e = po.fill_synthetic(BB, 77, 0, 1 << 10)
        ch = po.fill_synthetic(BB, 78, 0, 10)
        pi = zlib.SumcheckProver.prove_interactive(zlib.Multilinear.init(ctx, e), ch)
        wi = po.sumcheck_prove_interactive(BB, e, ch)
        assert pi.round_polynomials.tolist() == wi.round_polys.tolist() and pi.final_eval == wi.final_eval
    finally:
        zlib.lib().zh_set_grid_min_log2(old)
        ctx.set_option("prod_host_tail_log2", old_ht)


@pytest.mark.parametrize("host_tail", [1, 2, 4, 7, 10, 12])
def test_small_tables_finish_on_the_host(zlib, ctx, po, host_tail):
    """zb_prod_fold_dump + the host twin's own rounds (tables of <= 2^host_tail entries leave the device): every threshold, table
    sizes on both sides of it, all degrees, consuming and not — proofs equal the oracle's bit for bit, a consumed table ends as
    its final evaluation, an untouched one stays untouched."""
    old_ht = ctx.get_option("prod_host_tail_log2")
    ctx.set_option("prod_host_tail_log2", host_tail)
    try:
        for d in (1, 2, 3):
            for lg in sorted({1, 2, host_tail, host_tail + 1, host_tail + 2, host_tail + 3, 12, 15}):
                es = [po.fill_synthetic(BB, 4321 + 7 * k + lg, 0, 1 << lg) for k in range(d)]
                want = po.prodcheck_prove(BB, es)
                for consume in (False, True):
                    polys = [zlib.Multilinear.init(ctx, e) for e in es]
                    pr = zlib.ProductSumcheckProver.prove(polys, consume=consume)
                    assert pr.claimed_sum == want.claimed_sum, (d, lg)
                    assert pr.round_polynomials.tolist() == want.round_polys.tolist(), (d, lg, consume)
                    assert pr.final_point.tolist() == want.final_point.tolist() and pr.final_evals == want.final_evals
                    for k in range(d):
                        if consume:
                            assert polys[k].evaluations.tolist() == [want.final_evals[k]], (d, lg, k)
                        else:
                            assert np.array_equal(polys[k].evaluations, es[k])
                        polys[k].deinit()
        e = po.fill_synthetic(BB, 77, 0, 1 << 11)
        ch = po.fill_synthetic(BB, 78, 0, 11)
        ctx.set_option("linear_d1", 0)  # proveInteractive through the product path (d = 1) with fixed challenges
        pi = zlib.SumcheckProver.prove_interactive(zlib.Multilinear.init(ctx, e), ch)
        wi = po.sumcheck_prove_interactive(BB, e, ch)
        assert pi.round_polynomials.tolist() == wi.round_polys.tolist() and pi.final_eval == wi.final_eval
    finally:
        ctx.set_option("linear_d1", 1)
        ctx.set_option("prod_host_tail_log2", old_ht)


def test_fold_dump_entry_directly(zlib, ctx, po):
    """zb_prod_fold_dump against partialEval applied nfold times (multilinear.zig:154-180); error behaviour."""
    import ctypes as C
    L = zlib.lib()
    for d in (1, 2, 3):
        for lg, nfold in ((0, 0), (1, 1), (2, 2), (5, 0), (9, 1), (12, 2), (10, 0), (11, 1), (12, 0), (13, 1), (14, 2)):
            es = [po.fill_synthetic(BB, 900 + k + lg, 0, 1 << lg) for k in range(d)]
            polys = [zlib.Multilinear.init(ctx, e) for e in es]
            hs = (C.c_uint64 * d)(*[p.handle for p in polys])
            r = [int(x) for x in po.fill_synthetic(BB, 55 + lg, 0, 2)]
            rr = (C.c_uint64 * 2)(*r)
            m = (1 << lg) >> nfold
            out = (C.c_uint32 * (d * m))()
            ctx.check(L.zb_prod_fold_dump(ctx.handle, hs, d, nfold, rr, out))
            got = np.frombuffer(out, dtype=np.uint32).reshape(d, m)
            for k in range(d):
                w = np.array(es[k], dtype=np.uint64)
                for t in range(nfold):
                    w = po.mle_partial_eval(BB, w, r[t])
                assert got[k].tolist() == [int(x) for x in w], (d, lg, nfold, k)
                assert np.array_equal(polys[k].evaluations, es[k])  # the device tables are not modified
    big = [zlib.Multilinear.init(ctx, po.fill_synthetic(BB, 1, 0, 1 << 14))]
    hs = (C.c_uint64 * 1)(big[0].handle)
    out = (C.c_uint32 * 8192)()
    rr = (C.c_uint64 * 2)(1, BB)
    assert L.zb_prod_fold_dump(ctx.handle, hs, 1, 1, rr, out) == -22  # BadArgument: 2^13 folded entries are too many
    assert L.zb_prod_fold_dump(ctx.handle, hs, 1, 2, rr, out) == -20  # NotCanonical
    assert L.zb_prod_fold_dump(ctx.handle, hs, 1, 3, rr, out) == -22


def test_grid_entry_points_directly(zlib, ctx, po):
    """zb_prod_grid / zb_prod_fold_grid against the oracle's single-round functions: g(X) = G(X,0)+G(X,1), g'(Y) = G(r,Y)."""
    import ctypes as C
    L = zlib.lib()
    for d in (1, 2, 3):
        np_ = 2 if d == 1 else d + 1
        es = [po.fill_synthetic(BB, 300 + k, 0, 1 << 9) for k in range(d)]
        polys = [zlib.Multilinear.init(ctx, e) for e in es]
        hs = (C.c_uint64 * d)(*[p.handle for p in polys])
        grid = (C.c_uint64 * 16)()
        ctx.check(L.zb_prod_grid(ctx.handle, hs, d, grid))
        g = np.array(list(grid)[:np_ * np_], dtype=object).reshape(np_, np_)
        want0 = po.prod_round_coeffs(BB, es)

        def to_coeffs(ev):
            ev = [int(x) for x in ev]
            inv2 = (BB + 1) // 2
            if d == 1:
                return [ev[0], (ev[1] - ev[0]) % BB]
            if d == 2:
                return [ev[0], (ev[1] - ev[0] - ev[2]) % BB, ev[2]]
            return [ev[0], ((ev[1] - ev[2]) * inv2 - ev[3]) % BB, ((ev[1] + ev[2]) * inv2 - ev[0]) % BB, ev[3]]

        def horner(c, x):
            acc = 0
            for a in reversed(c):
                acc = (acc * x + a) % BB
            return acc

        assert to_coeffs([(g[ix][0] + g[ix][1]) % BB for ix in range(np_)]) == want0
        r1, r2 = 123456789, 987654321
        folded = [po.mle_partial_eval(BB, e, r1) for e in es]
        want1 = po.prod_round_coeffs(BB, folded)
        assert to_coeffs([horner(to_coeffs([g[ix][iy] for ix in range(np_)]), r1) for iy in range(np_)]) == want1
        # fold both variables, in place, and compare tables + next grid's first round
        rr = (C.c_uint64 * 2)(r1, r2)
        ctx.check(L.zb_prod_fold_grid(ctx.handle, hs, d, 2, rr, None, grid))
        folded2 = [po.mle_partial_eval(BB, f, r2) for f in folded]
        for p, f in zip(polys, folded2):
            assert np.array_equal(p.evaluations, f)
        g = np.array(list(grid)[:np_ * np_], dtype=object).reshape(np_, np_)
        assert to_coeffs([(g[ix][0] + g[ix][1]) % BB for ix in range(np_)]) == po.prod_round_coeffs(BB, folded2)
        # one variable, out of place
        out = (C.c_uint64 * d)()
        r3 = (C.c_uint64 * 2)(55555, 0)
        ctx.check(L.zb_prod_fold_grid(ctx.handle, hs, d, 1, r3, out, grid))
        folded3 = [po.mle_partial_eval(BB, f, 55555) for f in folded2]
        for k in range(d):
            assert np.array_equal(zlib.Multilinear(ctx, out[k]).evaluations, folded3[k])
            assert np.array_equal(polys[k].evaluations, folded2[k])  # inputs untouched
        with pytest.raises(zlib.ZigzError):  # too small for the vector kernel: the caller must use the single-round entries
            small = zlib.Multilinear.init(ctx, [1, 2, 3, 4])
            ctx.check(L.zb_prod_grid(ctx.handle, (C.c_uint64 * 1)(small.handle), 1, grid))


# ---------------------------------------------------------------- d = 1: several rounds per pass (zb_mle_block_sums / zb_mle_fold_multi)
@pytest.mark.parametrize("lg", [11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22])
@pytest.mark.parametrize("host_tail", [12, 10, 7, 2])
def test_linear_prover_matches_one_round_per_kernel_and_oracle(zlib, ctx, po, lg, host_tail):
    """SumcheckProver.prove (sumcheck_prover.zig:26-91) through block sums + multi-variable folds: every pass shape
    (small tables: ONE block-sum pass of up to 2^8 blocks + ONE fold of up to 8 variables; large ones: passes of 5; published
    tables of 2^2..2^12 entries, left-over variable counts) against the oracle, the interactive variant, and the consuming
    variant (the polynomial ends as its final evaluation)."""
    e = po.fill_synthetic(BB, 0xC0FFEE + lg, 0, 1 << lg)
    want = po.sumcheck_prove(BB, e)
    ctx.set_option("host_tail_log2", host_tail)
    try:
        poly = zlib.Multilinear.init(ctx, e)
        assert zlib.SumcheckProver.prove(poly).to_bytes() == want.to_bytes()
        assert np.array_equal(poly.evaluations[:33], e[:33]) and len(poly) == 1 << lg  # untouched
        ch = po.fill_synthetic(BB, 99, 0, lg)
        wi = po.sumcheck_prove_interactive(BB, e, ch)
        pi = zlib.SumcheckProver.prove_interactive(poly, ch)
        assert pi.round_polynomials.tolist() == wi.round_polys.tolist() and pi.final_eval == wi.final_eval
        pc = zlib.ProductSumcheckProver.prove([poly], consume=True)
        assert pc.round_polynomials.tolist() == want.round_polys.tolist() and pc.final_evals[0] == want.final_eval
        assert len(poly) == 1 and int(poly.evaluations[0]) == want.final_eval
    finally:
        ctx.set_option("host_tail_log2", 12)


def test_block_sums_and_fold_multi_entry_points(zlib, ctx, po):
    """zb_mle_block_sums / zb_mle_fold_multi against partialEval applied k times (multilinear.zig:154-180) and plain block sums."""
    import ctypes as C
    L = zlib.lib()
    rng = np.random.default_rng(11)
    # (k_next == log2 of the folded length: the folded table itself comes back — the wide kernel, k up to 8)
    for lg, k, k_next in ((7, 1, 2), (9, 2, 5), (12, 3, 4), (13, 5, 5), (14, 4, 10), (12, 5, 7), (16, 5, 1), (8, 5, 3), (3, 1, 2), (10, 6, 4),
                          (14, 7, 7), (20, 8, 12), (16, 8, 8), (13, 1, 12), (10, 8, 2)):
        e = po.fill_synthetic(BB, 7000 + lg, 0, 1 << lg)
        poly = zlib.Multilinear.init(ctx, e)
        kk = k
        sums = np.zeros(1 << kk, np.uint64)
        ctx.check(L.zb_mle_block_sums(ctx.handle, poly.handle, kk, sums.ctypes.data_as(zlib.api.P64)))
        blocks = e.reshape(1 << kk, -1)
        assert sums.tolist() == [int(b.astype(object).sum() % BB) for b in blocks]
        r = rng.integers(0, BB, size=k, dtype=np.uint64)
        folded = e
        for x in r:
            folded = po.mle_partial_eval(BB, folded, int(x))
        for inplace in (False, True):
            out = C.c_uint64(0)
            got = np.zeros(1 << k_next, np.uint64)
            ctx.check(L.zb_mle_fold_multi(ctx.handle, poly.handle, k, r.ctypes.data_as(zlib.api.P64), None if inplace else C.byref(out),
                                          k_next, got.ctypes.data_as(zlib.api.P64)))
            res = poly if inplace else zlib.Multilinear(ctx, out.value)
            assert np.array_equal(res.evaluations, folded), (lg, k, inplace)
            assert got.tolist() == [int(b.astype(object).sum() % BB) for b in folded.reshape(1 << k_next, -1)]
            if not inplace:
                assert np.array_equal(poly.evaluations, e)
                res.deinit()
        poly.deinit()
    small = zlib.Multilinear.init(ctx, [1, 2, 3, 4, 5, 6, 7, 8])
    buf = np.zeros(32, np.uint64)
    assert L.zb_mle_block_sums(ctx.handle, small.handle, 2, buf.ctypes.data_as(zlib.api.P64)) == -22  # n < 4 * 2^k: BadArgument
    assert L.zb_mle_fold_multi(ctx.handle, small.handle, 1, buf.ctypes.data_as(zlib.api.P64), None, 1, buf.ctypes.data_as(zlib.api.P64)) == -22


# ---------------------------------------------------------------- extension: eq-weighted product sumcheck
def _eq_table(tau):
    """E[i] = prod_k (bit_k(i) ? tau[k] : 1 - tau[k]) — independent numpy/Python-int construction."""
    e = [1]
    for t in tau:  # bit k is added as the new TOP bit of the indices built so far
        t = int(t)
        e = [x * (1 - t) % BB for x in e] + [x * t % BB for x in e]
    return np.array(e, np.uint64)


@pytest.mark.parametrize("d", [1, 2])
@pytest.mark.parametrize("lg", [1, 2, 3, 6, 11, 16, 18])
def test_eq_weighted_sumcheck_vs_oracle(zlib, ctx, po, d, lg):
    """sum_x eq(tau, x) prod_k A_k(x) proved as the product sumcheck of (eq(tau, .), A_0, ..): same round polynomials as the
    oracle's product prover on the explicitly built eq table; for d = 1 the claimed sum is Multilinear.eval(tau); the final
    eq value is eq(tau, r) with the challenges in index-bit order."""
    n = 1 << lg
    tau = po.fill_synthetic(BB, 31337 + lg, 0, lg)
    es = [po.fill_synthetic(BB, 600 + 3 * lg + k, 0, n) for k in range(d)]
    polys = [zlib.Multilinear.init(ctx, e) for e in es]
    E = _eq_table(tau)
    import ctypes as C
    h = C.c_uint64(0)
    ctx.check(zlib.lib().zb_mle_eq(ctx.handle, tau.ctypes.data_as(zlib.api.P64), lg, C.byref(h)))
    eqp = zlib.Multilinear(ctx, h.value)
    assert np.array_equal(eqp.evaluations, E)
    eqp.deinit()
    want = po.prodcheck_prove(BB, [E] + es)
    pr = zlib.EqProductSumcheckProver.prove(tau, polys)
    assert pr.claimed_sum == want.claimed_sum
    assert pr.round_polynomials.tolist() == want.round_polys.tolist()
    assert pr.final_point.tolist() == want.final_point.tolist()
    assert pr.final_evals == want.final_evals
    if d == 1:
        assert pr.claimed_sum == polys[0].eval(tau) == po.mle_eval(BB, es[0], tau)
    r_by_bit = [int(x) for x in pr.final_point][::-1]  # round j binds index bit lg-1-j
    eq_at_r = 1
    for t, r in zip((int(x) for x in tau), r_by_bit):
        eq_at_r = eq_at_r * ((t * r + (1 - t) * (1 - r)) % BB) % BB
    assert pr.final_evals[0] == eq_at_r
    ok, final_claim = po.sumcheck_verify_rounds(BB, pr.round_polynomials, pr.claimed_sum)
    assert ok and final_claim == pr.final_eval
    assert np.array_equal(polys[0].evaluations, es[0])  # inputs untouched
    with pytest.raises(zlib.ZigzError):
        zlib.EqProductSumcheckProver.prove(tau[:-1] if lg > 1 else np.zeros(3, np.uint64), polys)
