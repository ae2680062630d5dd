"""Bit-for-bit comparison with the oracle AT THE SIZES BASELINE.json names (not only through properties):
C2 Lasso with 2^22 lookups per 8-bit table (and the non-power-of-two 3*2^20), the product sumcheck at 2^24 and 2^26
entries per table (the cp.async / bulk-copy ring kernels only engage from 2^16 entries), the reference's own d=1 prover
at 2^24/2^26, and a 2^24-leaf Merkle tree with 16 openings. The oracle needs seconds for each (≈ 1 min for the file)."""
import numpy as np
import pytest

from _cases import BB, lasso_queries, lasso_queries_np

pytestmark = pytest.mark.gpu


def test_vectorised_query_generator_matches_the_scalar_one():
    for op in ("add", "xor", "and"):
        assert np.array_equal(lasso_queries_np(op, 8, 300), lasso_queries(op, 8, 300))


def _same_lasso(got, want):
    sp = got.sumcheck_proof
    assert sp.round_polynomials.tolist() == want.sumcheck.round_polys.tolist()
    assert sp.final_point.tolist() == want.sumcheck.final_point.tolist()
    assert sp.final_eval == want.sumcheck.final_eval
    assert got.query_commitment == want.query_commitment
    assert got.table_commitment == want.table_commitment
    assert got.num_lookups == want.num_lookups


@pytest.mark.parametrize("op,nq", [("add", 1 << 22), ("and", 1 << 22), ("xor", 1 << 22), ("xor", 3 << 20)])
def test_c2_lasso_full_size_vs_oracle(zlib, ctx, po, op, nq):
    """BASELINE config C2 (lasso_prover.zig:103-173): 2^22 lookups into the 8-bit ADD/AND/XOR subtables; 3*2^20 lookups
    exercise the zero padding to 2^22 (:140-142). Whole LassoProof compared: sumcheck proof + both commitments."""
    code = {"add": po.TABLE_ADD, "xor": po.TABLE_XOR, "and": po.TABLE_AND}[op]
    q = lasso_queries_np(op, 8, nq)
    want = po.lasso_prove(BB, po.build_table(BB, code, 8), q)
    _same_lasso(zlib.LassoProver.prove_builtin(ctx, code, 8, q), want)


@pytest.mark.parametrize("d,lg", [(3, 24), (3, 26), (1, 24), (1, 26), (2, 25)])
def test_sumcheck_full_size_vs_oracle(zlib, ctx, po, d, lg):
    """sumcheck_prover.zig:26-91 (d = 1) and its degree-d product extension at 2^24-2^26 entries per table: every
    round polynomial, challenge and final evaluation against the oracle; d = 3 at 2^26 is the per-GPU shard size of
    BASELINE config C5 on 16 GPUs and runs every large-table kernel of the 2^30 job (grid passes with bulk-copy rings,
    one-round kernels, persistent tail)."""
    n = 1 << lg
    es = [po.fill_synthetic(BB, 0x5A49475A + k, 0, n) for k in range(d)]
    polys = [zlib.Multilinear.synthetic(ctx, 0x5A49475A + k, n) for k in range(d)]
    want = po.prodcheck_prove(BB, es)
    pr = zlib.ProductSumcheckProver.prove(polys, consume=False)
    assert pr.claimed_sum == want.claimed_sum
    assert pr.round_polynomials.tolist() == want.round_polys.tolist()
    assert pr.final_point.tolist() == want.final_point.tolist()
    assert pr.final_evals == want.final_evals
    pr2 = zlib.ProductSumcheckProver.prove(polys, consume=True)
    assert pr2.round_polynomials.tolist() == want.round_polys.tolist() and pr2.final_evals == want.final_evals
    if d == 1:
        sp = zlib.SumcheckProver.prove(zlib.Multilinear.synthetic(ctx, 0x5A49475A, n))
        assert sp.to_bytes() == po.sumcheck_prove(BB, es[0]).to_bytes()
    for p in polys:
        p.deinit()


def test_c3_merkle_2p24_root_and_16_paths_vs_oracle(zlib, ctx, po):
    """merkle_tree.zig:283-360 at 2^24 leaves: root, height and 16 opening paths bit-for-bit (2^25 Keccak-f on the CPU)."""
    lg = 24
    e = po.fill_synthetic(BB, 0x5A49475A + 9, 0, 1 << lg)
    poly = zlib.Multilinear.synthetic(ctx, 0x5A49475A + 9, 1 << lg)
    want = po.merkle_build(e)
    com, tree = zlib.CommitmentScheme.commit(poly)
    assert com.commitment == want.root and tree.height == want.height == lg
    rng = np.random.default_rng(24)
    indices = [0, (1 << lg) - 1] + [int(x) for x in rng.integers(0, 1 << lg, size=14)]
    for idx, (v, sib, dirs) in zip(indices, po.merkle_open_many(want, indices)):
        pr = tree.open(idx)
        assert pr.value == v == int(e[idx])
        assert np.array_equal(pr.path.siblings, sib) and np.array_equal(pr.path.directions, dirs)
    tree.deinit()
    poly.deinit()


def test_tree_keeps_its_values_when_the_polynomial_is_consumed(zlib, ctx, po):
    """SimpleMerkleTree.build dupes the values (merkle_tree.zig:291): commit, then a CONSUMING sumcheck that folds the
    polynomial in place, then open: the opening must still report the committed leaf and verify."""
    for lg in (3, 12, 17):
        e = po.fill_synthetic(BB, 4242 + lg, 0, 1 << lg)
        poly = zlib.Multilinear.init(ctx, e)
        com, tree = zlib.CommitmentScheme.commit(poly)
        zlib.ProductSumcheckProver.prove([poly], consume=True)
        assert len(poly) == 1
        want = po.merkle_build(e)
        for idx in (0, 1, (1 << lg) - 1):
            pr = tree.open(idx)
            assert pr.value == int(e[idx])
            assert po.merkle_verify(want.root, pr.value, pr.path.siblings, pr.path.directions)
            assert zlib.SimpleMerkleTree.verify(com.commitment, pr)
        tree.deinit()
