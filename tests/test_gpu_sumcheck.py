"""SumcheckProver on the GPU (device rounds + host transcript) against the golden vectors and the oracle."""
import hashlib

import numpy as np
import pytest

from _cases import BB, assert_sumcheck_equal, sumcheck_case_evals, synthetic

pytestmark = pytest.mark.gpu


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
