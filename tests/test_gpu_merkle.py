"""SHA3-256 Merkle commitment kernels + CommitmentScheme against golden vectors, the oracle and the host verifier."""
import numpy as np
import pytest

from _cases import BB, synthetic

pytestmark = pytest.mark.gpu


def _vals(case):
    return np.array(case["values"], np.uint64) if case["values"] else synthetic(case["seed"], case["n"])


def test_golden_trees(zlib, ctx, golden):
    for name, case in golden["merkle"].items():
        t = zlib.SimpleMerkleTree.build(ctx, _vals(case))
        assert t.get_root().hex() == case["root"], name
        assert t.height == case["height"]
        assert t.leaf_hashes()[0].tobytes().hex() == case["leaf0"]
        for idx, o in case["opens"].items():
            pr = t.open(int(idx))
            assert pr.value == o["value"] and pr.path.directions.tolist() == o["dirs"]
            assert [s.tobytes().hex() for s in pr.path.siblings] == o["siblings"]
            assert zlib.SimpleMerkleTree.verify(t.get_root(), pr)
        t.deinit()


def test_reference_unit_test_facts(zlib, ctx):  # src/commitments/merkle_tree.zig:425-571
    t = zlib.SimpleMerkleTree.build(ctx, [1, 2, 3, 4])
    assert t.height == 2
    assert zlib.SimpleMerkleTree.build(ctx, [1, 2, 3, 4]).get_root() == t.get_root()  # determinism :437-451
    assert zlib.SimpleMerkleTree.build(ctx, [1, 2, 3, 5]).get_root() != t.get_root()
    pr = t.open(2)  # :453-468
    assert pr.value == 3 and pr.index == 2 and len(pr.path.siblings) == 2
    assert zlib.SimpleMerkleTree.verify(t.get_root(), pr)  # :470-488
    bad = zlib.MerkleOpeningProof(99, 2, pr.path)  # :490-508
    assert not zlib.SimpleMerkleTree.verify(t.get_root(), bad)
    t5 = zlib.SimpleMerkleTree.build(ctx, [1, 2, 3, 4, 5])  # :510-529
    assert t5.height == 3
    for i in range(5):
        assert zlib.SimpleMerkleTree.verify(t5.get_root(), t5.open(i))
    t1 = zlib.SimpleMerkleTree.build(ctx, [42])  # :531-545
    assert t1.height == 0 and len(t1.open(0).path.siblings) == 0
    assert zlib.SimpleMerkleTree.verify(t1.get_root(), t1.open(0))
    for lg in range(2, 7):  # :547-571
        tt = zlib.SimpleMerkleTree.build(ctx, np.arange(1, (1 << lg) + 1, dtype=np.uint64))
        assert len(tt.open(0).path.siblings) == lg
    with pytest.raises(zlib.ZigzError) as e:
        zlib.SimpleMerkleTree.build(ctx, np.zeros(0, np.uint64))
    assert e.value.name == "EmptyValues"  # :284
    with pytest.raises(zlib.ZigzError) as e:
        t5.open(5)
    assert e.value.name == "IndexOutOfBounds"  # :325 (index >= values.len, NOT the padded size)


def test_pointer_variant_alias(zlib, ctx, po):  # merkle_tree.zig:78-264: same padding, same root; open is NotImplemented
    for n in (1, 4, 5, 100):
        vals = synthetic(n + 9, n)
        t = zlib.MerkleTree.build(ctx, vals)
        assert t.root_hash() == po.merkle_build(vals).root
        pr = zlib.SimpleMerkleTree.build(ctx, vals).open(n - 1)
        assert zlib.MerkleTree.verify(t.root_hash(), pr)
        with pytest.raises(zlib.ZigzError):
            t.open(0)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 100, 1023, 1024, 1025, 2048, 5000, 1 << 14])
def test_build_and_open_vs_oracle(zlib, ctx, po, n):
    vals = synthetic(n, n)
    if n > 2:
        vals[1], vals[2] = BB - 1, 0
    want = po.merkle_build(vals)
    t = zlib.SimpleMerkleTree.build(ctx, vals)
    assert t.get_root() == want.root and t.height == want.height
    assert np.array_equal(t.leaf_hashes(), want.leaf_hashes)
    rng = np.random.default_rng(n)
    for idx in {0, n - 1, *[int(x) for x in rng.integers(0, n, size=4)]}:
        pr = t.open(idx)
        v, sib, dirs = po.merkle_open(want, idx)
        assert pr.value == v
        assert np.array_equal(pr.path.siblings, sib) and np.array_equal(pr.path.directions, dirs)
        assert po.merkle_verify(want.root, pr.value, pr.path.siblings, pr.path.directions)
    t.deinit()


def test_commitment_scheme_vs_oracle(zlib, ctx, po):  # src/commitments/polynomial_commit.zig:261-455
    for lg in (0, 1, 3, 10):
        e = synthetic(31 + lg, 1 << lg)
        poly = zlib.Multilinear.init(ctx, e)
        com, tree = zlib.CommitmentScheme.commit(poly)
        want = po.merkle_build(e)
        assert com.commitment == want.root and com.num_vars == lg
        pt = synthetic(77, lg)
        op = zlib.CommitmentScheme.open(poly, tree, pt)
        value, li, lv, sib, dirs = po.commit_open(BB, want, pt)
        assert (op.value, op.merkle_proof.index, op.merkle_proof.value) == (value, li, lv)
        assert np.array_equal(op.merkle_proof.path.siblings, sib) and np.array_equal(op.merkle_proof.path.directions, dirs)
        assert zlib.CommitmentScheme.verify(com, op)
        if lg:
            with pytest.raises(zlib.ZigzError) as err:
                zlib.CommitmentScheme.open(poly, tree, pt[:-1])
            assert err.value.name == "PointDimensionMismatch"  # :92-94
            tampered = zlib.OpeningProof(op.point, op.value, zlib.MerkleOpeningProof((lv + 1) % BB, li, op.merkle_proof.path))
            assert not zlib.CommitmentScheme.verify(com, tampered)  # :322-343


def test_batch_commit_43_witness_polynomials(zlib, ctx, po):
    """Prover.generateCommitments commits 43 polynomials of equal length (src/prover/prover.zig:405-410)."""
    lg = 8
    es = [synthetic(500 + i, 1 << lg) for i in range(43)]
    polys = [zlib.Multilinear.init(ctx, e) for e in es]
    coms, trees = zlib.CommitmentScheme.batch_commit(polys)
    assert len(coms) == 43
    for e, c, t in zip(es, coms, trees):
        assert c.commitment == po.merkle_build(e).root
    proofs = [zlib.CommitmentScheme.open(p, t, synthetic(i, lg)) for i, (p, t) in enumerate(zip(polys, trees))]
    assert zlib.CommitmentScheme.batch_verify(coms, proofs)  # :376-411
    assert not zlib.CommitmentScheme.batch_verify(coms[:-1], proofs)


def test_config_c3_size_tree(zlib, ctx, po):
    """BASELINE config C3: 2^26-entry witness polynomial. Openings verify on the host; the root equals the hash of the
    roots of the two half-size trees (the identity subtree sharding relies on)."""
    lg = 26
    poly = zlib.Multilinear.synthetic(ctx, 0xBEEF, 1 << lg)
    com, tree = zlib.CommitmentScheme.commit(poly)
    rng = np.random.default_rng(2)
    from _cases import splitmix64
    for idx in [0, (1 << lg) - 1] + [int(x) for x in rng.integers(0, 1 << lg, size=6)]:
        pr = tree.open(idx)
        assert pr.value == splitmix64(0xBEEF + idx) % BB and len(pr.path.siblings) == lg
        assert zlib.SimpleMerkleTree.verify(com.commitment, pr)
    tree.deinit()
    lo = zlib.Multilinear.synthetic(ctx, 0xBEEF, 1 << (lg - 1))
    hi = zlib.Multilinear.synthetic(ctx, 0xBEEF, 1 << (lg - 1), start=1 << (lg - 1))
    (clo, chi), trees = zlib.CommitmentScheme.batch_commit([lo, hi])
    assert zlib.sha3_256(clo.commitment + chi.commitment) == com.commitment
    for t in trees:
        t.deinit()
    for m in (poly, lo, hi):
        m.deinit()


@pytest.mark.parametrize("lg", [22, 26])
def test_full_size_tree_properties(zlib, ctx, po, lg):
    """2^22 and 2^26 leaves (BASELINE config C3): every opened path must verify against the root with the HOST verifier,
    leaf digests must equal SHA3(le64(value)), and the root must equal the root of the two half-trees hashed together
    (subtree sharding)."""
    poly = zlib.Multilinear.synthetic(ctx, 0xC0FFEE, 1 << lg)
    com, tree = zlib.CommitmentScheme.commit(poly)
    from _cases import splitmix64
    rng = np.random.default_rng(1)
    for idx in [0, (1 << lg) - 1] + [int(x) for x in rng.integers(0, 1 << lg, size=14)]:
        pr = tree.open(idx)
        assert pr.value == splitmix64(0xC0FFEE + idx) % BB
        assert len(pr.path.siblings) == lg
        assert zlib.SimpleMerkleTree.verify(com.commitment, pr)
        assert po.merkle_verify(com.commitment, pr.value, pr.path.siblings, pr.path.directions)
    ev = poly.evaluations
    half = 1 << (lg - 1)
    lo = zlib.SimpleMerkleTree.build(ctx, ev[:half])
    hi = zlib.SimpleMerkleTree.build(ctx, ev[half:])
    assert zlib.sha3_256(lo.get_root() + hi.get_root()) == com.commitment
    for t in (tree, lo, hi):
        t.deinit()
    poly.deinit()
