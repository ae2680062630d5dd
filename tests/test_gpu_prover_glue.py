"""Prover.generateCommitments (batched commit + transcript interleave + openings) and witness packing on the GPU."""
import hashlib

import numpy as np
import pytest

from _cases import BB, synthetic, witness_cols

pytestmark = pytest.mark.gpu


def _check(c, want_roots, want_open):
    assert [r.tobytes().hex() for r in c.roots] == want_roots
    for i, o in enumerate(want_open):
        assert c.points[i].tolist() == o["point"] and int(c.values[i]) == o["value"], i
        assert (int(c.leaf_indices[i]), int(c.leaf_values[i])) == (o["leaf_index"], o["leaf_value"])
        assert [s.tobytes().hex() for s in c.siblings[i]] == o["siblings"] and c.dirs[i].tolist() == o["dirs"]


def test_generate_commitments_golden(zlib, ctx, golden):
    for name, case in golden["generate_commitments"].items():
        polys = [zlib.Multilinear.init(ctx, synthetic(case["seed"] + i, 1 << case["lg"])) for i in range(case["count"])]
        tr = zlib.FiatShamirTranscript()
        tr.append_bytes(b"PROGRAM")
        tr.append_field_element(4096)
        c = zlib.generate_commitments(tr, polys)
        _check(c, case["roots"], case["openings"])
        assert tr.challenge() == case["next_challenge"], name  # transcript advanced exactly like the reference


@pytest.mark.parametrize("lg", [1, 5, 10, 13])
def test_generate_commitments_43_polys_vs_oracle(zlib, ctx, po, lg):
    es = [synthetic(40 + i, 1 << lg) for i in range(43)]
    polys = [zlib.Multilinear.init(ctx, e) for e in es]
    tr, otr = zlib.FiatShamirTranscript(), po.Transcript()
    for t in (tr, otr):
        t.append_bytes(b"LASSO_BEGIN")
    c = zlib.generate_commitments(tr, polys)
    w = po.generate_commitments(BB, otr, es)
    assert np.array_equal(c.roots, w.roots) and np.array_equal(c.points, w.points)
    assert np.array_equal(c.values, w.values) and np.array_equal(c.leaf_indices, w.leaf_indices)
    assert np.array_equal(c.leaf_values, w.leaf_values)
    assert np.array_equal(c.siblings, w.siblings) and np.array_equal(c.dirs, w.dirs)
    assert tr.challenge() == otr.challenge(BB)
    # every opening verifies against its commitment (Verifier.verifyOpening, src/verifier/verifier.zig:270-294)
    for i in range(43):
        proof = zlib.MerkleOpeningProof(int(c.leaf_values[i]), int(c.leaf_indices[i]), zlib.MerklePath(c.siblings[i], c.dirs[i]))
        assert zlib.SimpleMerkleTree.verify(c.roots[i].tobytes(), proof)


def test_witness_pack_golden_and_oracle(zlib, ctx, po, golden):
    for name, case in golden["witness_pack"].items():
        polys = zlib.witness_pack(ctx, witness_cols(case["steps"]))
        out = np.stack([p.evaluations for p in polys])
        assert hashlib.sha3_256(out.astype("<u8").tobytes()).hexdigest() == case["packed_sha3"], name
    for steps in (1, 2, 3, 64, 65, 1000, 4097):
        cols = witness_cols(steps)
        cols[5, -1] = 2**64 - 1  # F.init reduces any u64
        polys = zlib.witness_pack(ctx, cols)
        want = po.witness_pack(BB, cols, 33)
        assert len(polys) == 43 and polys[0].num_vars == (steps - 1).bit_length()
        for p, w in zip(polys, want):
            assert np.array_equal(p.evaluations, w)


def test_trace_to_openings_pipeline(zlib, ctx, po):
    """witness packing -> generateCommitments on the packed polynomials, end to end against the oracle."""
    cols = witness_cols(300)
    polys = zlib.witness_pack(ctx, cols)
    es = list(po.witness_pack(BB, cols, 33))
    tr, otr = zlib.FiatShamirTranscript(), po.Transcript()
    c = zlib.generate_commitments(tr, polys)
    w = po.generate_commitments(BB, otr, es)
    assert np.array_equal(c.roots, w.roots) and np.array_equal(c.values, w.values) and np.array_equal(c.siblings, w.siblings)


def test_prove_from_trace_bytes_identical(zlib, ctx, po, golden):
    """`zigz prove` after the VM: proof bytes equal the golden digests and the oracle's bytes; the verifier accepts."""
    from _cases import prove_inputs
    for name, case in golden["prove_from_trace"].items():
        inp = prove_inputs(case["steps"], case["seed"], case["n_init"], case["n_out"])
        proof = zlib.prove_from_trace(ctx, **inp)
        assert len(proof) == case["proof_len"] and proof[:96].hex() == case["proof_head"]
        assert hashlib.sha3_256(proof).hexdigest() == case["proof_sha3"], name
        assert proof == po.prove_from_trace(BB, **inp)
        assert zlib.verify_proof(proof, inp["program"]) == "Accept" == po.verify_proof(BB, proof, inp["program"])
        assert zlib.prove_from_trace(ctx, compat_buffer=True, **inp) == proof
    # determinism of opening points across two provers, different entry pc => different points (integration_tests.zig:212-245)
    inp = prove_inputs(16, 21, 3, 0)
    a, b = zlib.prove_from_trace(ctx, **inp), zlib.prove_from_trace(ctx, **inp)
    assert a == b
    inp2 = dict(inp, entry_pc=0x2000)
    assert zlib.prove_from_trace(ctx, **inp2) != a
    with pytest.raises(zlib.ZigzError) as e:
        zlib.prove_from_trace(ctx, b"", 0, [], np.zeros((43, 0), np.uint64), 0, [0] * 32, [])
    assert e.value.name == "EmptyTrace"


def test_prove_from_trace_larger_trace_vs_oracle(zlib, ctx, po):
    from _cases import prove_inputs
    inp = prove_inputs(1000, 33, 32, 5)  # pads to 1024 steps, ~400 lookup constraints
    proof = zlib.prove_from_trace(ctx, **inp)
    assert proof == po.prove_from_trace(BB, **inp)
    assert zlib.verify_proof(proof, inp["program"]) == "Accept"
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 0x10
    assert zlib.verify_proof(bytes(bad), inp["program"]) == po.verify_proof(BB, bytes(bad), inp["program"])


@pytest.mark.parametrize("steps", [1, 2, 5, 64, 1000, 4097, (1 << 17) + 3])
def test_witness_pack_commit_pipeline_equals_the_two_calls(zlib, ctx, po, steps):
    """zb_witness_pack_commit (trace upload overlapped with leaf hashing) == zb_witness_pack followed by zb_merkle_build:
    same polynomials (witness.zig:29-270), same roots and openings (merkle_tree.zig:283-360); the large case crosses the
    pinned-staging threshold of the upload path."""
    if steps > 100000:
        rng = np.random.default_rng(steps)
        cols = rng.integers(0, 1 << 64, size=(43, steps), dtype=np.uint64)
    else:
        cols = witness_cols(steps)
    polys, coms, trees = zlib.witness_pack_commit(ctx, cols)
    ref = zlib.witness_pack(ctx, cols)
    rcoms, rtrees = zlib.CommitmentScheme.batch_commit(ref)
    assert len(polys) == 43
    for i in (0, 1, 32, 33, 42) if steps > 100000 else range(43):
        assert np.array_equal(polys[i].evaluations, ref[i].evaluations), i
    for a, b in zip(coms, rcoms):
        assert a.commitment == b.commitment and a.num_vars == b.num_vars
    if steps <= 1000:
        want = po.witness_pack(BB, cols, 33)
        for i in (0, 7, 33, 42):
            assert coms[i].commitment == po.merkle_build(want[i]).root
    padded = len(polys[0])
    for i in (0, 42):
        for idx in {0, padded - 1, padded // 2}:
            pa, pb = trees[i].open(idx), rtrees[i].open(idx)
            assert pa.value == pb.value and np.array_equal(pa.path.siblings, pb.path.siblings)
    for x in polys + ref:
        x.deinit()
    for t in trees + rtrees:
        t.deinit()
