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
