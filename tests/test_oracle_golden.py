"""The C oracle against the committed golden vectors (tests/golden/zigz_golden.json, produced by the independent
pure-Python restatement tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from _cases import BB, assert_sumcheck_equal, lasso_queries, sumcheck_case_evals, synthetic


def test_transcript(po, golden):
    t = po.Transcript()
    assert [t.challenge(BB) for _ in range(3)] == golden["transcript_first_challenges"]
    t = po.Transcript()
    t.append_bytes(b"SUMCHECK_BEGIN")
    t.append_field(12345)
    assert [t.challenge(BB), t.challenge(BB)] == golden["transcript_mixed"]
    assert po.hash_leaf(0).hex() == golden["sha3_leaf_0"]
    assert po.hash_leaf(BB - 1).hex() == golden["sha3_leaf_p_minus_1"]


def test_synthetic_generator_matches(po):
    assert po.fill_synthetic(BB, 0x5A49475A, 3, 50).tolist() == synthetic(0x5A49475A, 50, start=3).tolist()


def test_sumcheck(po, golden):
    for name, case in golden["sumcheck"].items():
        e = sumcheck_case_evals(case)
        if "challenges" in case:
            pr = po.sumcheck_prove_interactive(case["p"], e, case["challenges"])
        else:
            pr = po.sumcheck_prove(case["p"], e)
            assert pr.claimed_sum == case["claimed_sum"], name
        assert_sumcheck_equal(pr, case)
        if case["p"] == BB or True:
            assert hashlib.sha3_256(pr.to_bytes()).hexdigest() == case["to_bytes_sha3"], name


def test_prodcheck(po, golden):
    for name, case in golden["prodcheck"].items():
        polys = [synthetic(case["seed"] + k, case["n"]) for k in range(case["d"])]
        pr = po.prodcheck_prove(BB, polys)
        assert pr.claimed_sum == case["claimed_sum"], name
        assert_sumcheck_equal(pr, case)
    # d == 1 must be bit-identical to the reference sumcheck
    e = synthetic(0x5A49475A, 256)
    a, b = po.prodcheck_prove(BB, [e]), po.sumcheck_prove(BB, e)
    assert a.round_polys.tolist() == b.round_polys.tolist() and a.final_point.tolist() == b.final_point.tolist()
    assert a.final_evals[0] == b.final_eval


def test_eval(po, golden):
    for name, case in golden["eval"].items():
        if "seed" not in case:
            continue
        e = synthetic(case["seed"], case["n"])
        assert po.mle_eval(BB, e, case["point"]) == case["value"], name
    assert [po.mle_eval(17, [0, 1], [2]), po.mle_eval(17, [0, 1], [5])] == golden["eval"]["f17_x_at_2_5"]["values"]


def _merkle_values(case):
    return np.array(case["values"], np.uint64) if case["values"] else synthetic(case["seed"], case["n"])


def test_merkle(po, golden):
    for name, case in golden["merkle"].items():
        t = po.merkle_build(_merkle_values(case))
        assert t.root.hex() == case["root"], name
        assert t.height == case["height"]
        assert t.leaf_hashes[0].tobytes().hex() == case["leaf0"]
        for idx, o in case["opens"].items():
            v, sib, dirs = po.merkle_open(t, int(idx))
            assert v == o["value"] and dirs.tolist() == o["dirs"]
            assert [s.tobytes().hex() for s in sib] == o["siblings"]
            assert po.merkle_verify(t.root, v, sib, dirs)


def test_lasso(po, golden):
    g = golden["lasso"]
    assert po.lasso_hash_row(BB, [1, 2, 3]) == g["hash_entry_1_2_3"]
    assert po.lasso_commit_poly([1, 2, 3, 4]).hex() == g["flat_commit_1234"]
    xor2 = po.build_table(BB, po.TABLE_XOR, 2)
    c = g["xor2_queries"]
    for mapping in (None, c["mapping"]):
        pr = po.lasso_prove(BB, xor2, c["queries"], mapping=mapping)
        assert_sumcheck_equal(pr.sumcheck, c["sumcheck"])
        assert (pr.query_commitment.hex(), pr.table_commitment.hex()) == (c["query_commitment"], c["table_commitment"])
    for op, code in (("add", po.TABLE_ADD), ("xor", po.TABLE_XOR), ("and", po.TABLE_AND)):
        c = g[f"{op}4_200"]
        pr = po.lasso_prove(BB, po.build_table(BB, code, 4), lasso_queries(op, 4, 200))
        assert pr.sumcheck.num_vars == c["num_vars"] == 8
        assert_sumcheck_equal(pr.sumcheck, c["sumcheck"])
        assert (pr.query_commitment.hex(), pr.table_commitment.hex()) == (c["query_commitment"], c["table_commitment"])


def test_generate_commitments(po, golden):
    for name, case in golden["generate_commitments"].items():
        polys = [synthetic(case["seed"] + i, 1 << case["lg"]) for i in range(case["count"])]
        tr = po.Transcript()
        tr.append_bytes(b"PROGRAM")
        tr.append_field(4096)
        c = po.generate_commitments(BB, tr, polys)
        assert [r.tobytes().hex() for r in c.roots] == case["roots"], name
        for i, o in enumerate(case["openings"]):
            assert c.points[i].tolist() == o["point"] and int(c.values[i]) == o["value"]
            assert (int(c.leaf_indices[i]), int(c.leaf_values[i])) == (o["leaf_index"], o["leaf_value"])
            assert [s.tobytes().hex() for s in c.siblings[i]] == o["siblings"] and c.dirs[i].tolist() == o["dirs"]
        assert tr.challenge(BB) == case["next_challenge"]  # the transcript was advanced identically


def test_witness_pack(po, golden):
    import hashlib
    from _cases import witness_cols
    for name, case in golden["witness_pack"].items():
        out = po.witness_pack(BB, witness_cols(case["steps"]), 33)
        assert hashlib.sha3_256(out.astype("<u8").tobytes()).hexdigest() == case["packed_sha3"], name
        assert out[0].tolist() == case["first_col"] and out[42].tolist() == case["last_col"]


def test_prove_from_trace_and_verify(po, golden):
    import hashlib
    from _cases import prove_inputs
    for name, case in golden["prove_from_trace"].items():
        inp = prove_inputs(case["steps"], case["seed"], case["n_init"], case["n_out"])
        proof = po.prove_from_trace(BB, **inp)
        assert len(proof) == case["proof_len"], name
        assert proof[:96].hex() == case["proof_head"]
        assert hashlib.sha3_256(proof).hexdigest() == case["proof_sha3"], name
        assert po.verify_proof(BB, proof, inp["program"]) == "Accept"
        # integration-test behaviours (tests/integration_tests.zig:55-375)
        import pytest
        with pytest.raises(po.OracleError) as e:
            po.verify_proof(BB, proof, inp["program"] + b"\x00")
        assert e.value.name == "ProgramHashMismatch"
        if case["steps"] > 1:
            bad = bytearray(proof)
            bad[-40] ^= 1  # a sibling digest of the last opening
            assert po.verify_proof(BB, bytes(bad), inp["program"]) == "RejectInvalidCommitment"
    # the reference's fixed serialization buffer is under-estimated: large num_vars cannot be serialized
    import pytest
    inp = prove_inputs(64, 9, 0, 0)
    assert po.prove_from_trace(BB, compat_buffer=True, **inp) == po.prove_from_trace(BB, **inp)
    L = po.lib()
    assert L.zo_proof_exact_size(1 << 18, 0, 0, 0) <= L.zo_proof_estimated_size(1 << 18, 0, 0)
    assert L.zo_proof_exact_size(1 << 19, 0, 0, 0) > L.zo_proof_estimated_size(1 << 19, 0, 0)  # num_vars >= 19: NoSpaceLeft
    with pytest.raises(po.OracleError) as e:
        po.prove_from_trace(BB, b"", 0, [], np.zeros((43, 0), np.uint64), 0, [0] * 32, [])
    assert e.value.name == "EmptyTrace"
