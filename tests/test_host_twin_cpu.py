"""Host twin (C++ FiatShamirTranscript / SHA3 / field / verify / serialisation) against the oracle and the golden
vectors. These run on the CPU box: none of them touches the device ABI."""
import hashlib
import random

import numpy as np

from _cases import BB


def test_sha3_host_vs_hashlib(zlib):
    rng = random.Random(5)
    for n in list(range(0, 300)) + [136 * 5, 136 * 5 + 3, 8 * 17 * 9, 5000]:
        data = bytes(rng.getrandbits(8) for _ in range(n))
        assert zlib.sha3_256(data) == hashlib.sha3_256(data).digest(), n


def test_transcript_vs_golden_and_oracle(zlib, po, golden):
    t = zlib.FiatShamirTranscript()
    assert [t.challenge() for _ in range(3)] == golden["transcript_first_challenges"]
    t = zlib.FiatShamirTranscript()
    t.append_bytes(b"SUMCHECK_BEGIN")
    t.append_field_element(12345)
    assert [t.challenge(), t.challenge()] == golden["transcript_mixed"]
    # long mixed stream, misaligned appends, against the oracle
    rng = random.Random(9)
    a, b = zlib.FiatShamirTranscript(), po.Transcript()
    for step in range(400):
        k = rng.randrange(4)
        if k == 0:
            v = rng.randrange(BB)
            a.append_field_element(v)
            b.append_field(v)
        elif k == 1:
            data = bytes(rng.getrandbits(8) for _ in range(rng.randrange(0, 70)))
            a.append_bytes(data)
            b.append_bytes(data)
        elif k == 2:
            vs = [rng.randrange(BB) for _ in range(rng.randrange(1, 40))]
            a.append_field_elements(vs)
            for v in vs:
                b.append_field(v)
        else:
            assert a.challenge() == b.challenge(BB), step


def test_field_and_horner(zlib, po):
    L = zlib.lib()
    rng = random.Random(3)
    for _ in range(2000):
        x, y = rng.randrange(BB), rng.randrange(BB)
        assert L.zh_f_add(x, y) == po.lib().zo_f_add(BB, x, y)
        assert L.zh_f_sub(x, y) == po.lib().zo_f_sub(BB, x, y)
        assert L.zh_f_mul(x, y) == po.lib().zo_f_mul(BB, x, y)
    assert [zlib.eval_univariate_coeffs([3, 5], x) for x in (0, 1, 2)] == [3, 8, 13]  # sumcheck_protocol.zig:219-236
    for _ in range(200):
        c = [rng.randrange(BB) for _ in range(rng.randrange(1, 5))]
        x = rng.randrange(BB)
        assert zlib.eval_univariate_coeffs(c, x) == po.eval_univariate(BB, c, x)


def test_proof_to_bytes(zlib, po, golden):
    case = golden["sumcheck"]["bb_1to8"]
    pr = zlib.SumcheckProof(3, np.array(case["round_polys"], np.uint64), np.array(case["final_point"], np.uint64),
                            case["final_eval"])
    b = pr.to_bytes()
    assert len(b) == (2 + 3 * 3) * 8
    assert hashlib.sha3_256(b).hexdigest() == case["to_bytes_sha3"]


def test_merkle_verify_host(zlib, po, golden):
    for name, case in golden["merkle"].items():
        root = bytes.fromhex(case["root"])
        for idx, o in case["opens"].items():
            sib = np.array([list(bytes.fromhex(s)) for s in o["siblings"]], np.uint8).reshape(-1, 32)
            proof = zlib.MerkleOpeningProof(o["value"], int(idx), zlib.MerklePath(sib, np.array(o["dirs"], np.uint8)))
            assert zlib.SimpleMerkleTree.verify(root, proof), name
            bad = zlib.MerkleOpeningProof((o["value"] + 1) % BB, int(idx), proof.path)
            assert not zlib.SimpleMerkleTree.verify(root, bad)


def test_point_to_index(zlib):
    assert zlib.CommitmentScheme.point_to_index([]) == 0
    assert zlib.CommitmentScheme.point_to_index([13, 5, 6]) == 5
    assert zlib.CommitmentScheme.point_to_index([BB - 1] * 20) == (BB - 1) % (1 << 20)


def test_table_builders_match_oracle(zlib, po):
    for bits in (2, 4):
        assert np.array_equal(zlib.build_add_table(bits), po.build_table(BB, po.TABLE_ADD, bits))
        assert np.array_equal(zlib.build_xor_table(bits), po.build_table(BB, po.TABLE_XOR, bits))
        assert np.array_equal(zlib.build_and_table(bits), po.build_table(BB, po.TABLE_AND, bits))


def test_sha3_long_streams_hit_the_simd_absorb_loop(zlib):
    """Long word-aligned runs go through the AVX-512 absorb loop when the CPU has it; every split must agree with hashlib."""
    rng = random.Random(11)
    blob = bytes(rng.getrandbits(8) for _ in range(136 * 300 + 77))
    for n in (136 * 4 - 8, 136 * 4, 136 * 4 + 8, 136 * 5 + 64, 136 * 64, 136 * 257 + 40, len(blob)):
        assert zlib.sha3_256(blob[:n]) == hashlib.sha3_256(blob[:n]).digest(), n
    # transcript: partial block, then a long run of field elements, then bytes, then more elements
    for pre in (0, 1, 8, 64, 135, 136, 200):
        t = zlib.FiatShamirTranscript()
        h = hashlib.sha3_256()
        t.append_bytes(blob[:pre])
        h.update(blob[:pre])
        vals = [rng.randrange(BB) for _ in range(17 * 40 + 5)]
        t.append_field_elements(vals)
        h.update(b"".join(v.to_bytes(8, "little") for v in vals))
        t.append_bytes(blob[:3])
        h.update(blob[:3])
        t.append_field_elements(vals)
        h.update(b"".join(v.to_bytes(8, "little") for v in vals))
        assert t.finalize() == h.digest(), pre


def test_flat_commit_u64_and_u32_forms(zlib, po, golden):
    import ctypes as C
    L = zlib.lib()
    rng = np.random.default_rng(3)
    for n in (1, 4, 16, 17, 67, 68, 69, 17 * 9, 17 * 64 + 3, 50000):
        e = rng.integers(0, BB, size=n, dtype=np.uint64)
        want = hashlib.sha3_256(e.astype("<u8").tobytes()).digest()
        out = np.zeros(32, np.uint8)
        L.zh_flat_commit(e.ctypes.data_as(C.POINTER(C.c_uint64)), n, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert out.tobytes() == want == po.lasso_commit_poly(e), n
        e32 = e.astype(np.uint32)
        L.zh_flat_commit_u32(e32.ctypes.data_as(C.POINTER(C.c_uint32)), n, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert out.tobytes() == want, n
    out = np.zeros(32, np.uint8)
    e = np.array([1, 2, 3, 4], np.uint64)
    L.zh_flat_commit(e.ctypes.data_as(C.POINTER(C.c_uint64)), 4, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert out.tobytes().hex() == golden["lasso"]["flat_commit_1234"]


def test_verifier_twin_on_oracle_proofs(zlib, po, golden):
    """zh_verify_proof (host only) on proofs produced by the oracle: Accept, ProgramHashMismatch, tamper rejections
    (tests/integration_tests.zig:55-375 behaviours), and agreement with the oracle's verifier on random corruptions."""
    import pytest
    from _cases import prove_inputs
    rng = random.Random(4)
    for name, case in golden["prove_from_trace"].items():
        inp = prove_inputs(case["steps"], case["seed"], case["n_init"], case["n_out"])
        proof = po.prove_from_trace(BB, **inp)
        assert hashlib.sha3_256(proof).hexdigest() == case["proof_sha3"]
        assert zlib.verify_proof(proof, inp["program"]) == "Accept"
        with pytest.raises(zlib.ZigzError) as e:
            zlib.verify_proof(proof, inp["program"] + b"x")
        assert e.value.name == "ProgramHashMismatch"
        with pytest.raises(zlib.ZigzError) as e:
            zlib.verify_proof(b"ZIGX" + proof[4:], inp["program"])
        assert e.value.name == "InvalidProof"
        with pytest.raises(zlib.ZigzError):
            zlib.verify_proof(proof[:-5], inp["program"])
        for _ in range(60):  # single-byte corruptions: same verdict / error as the oracle's verifier
            bad = bytearray(proof)
            pos = rng.randrange(len(bad))
            bad[pos] ^= 1 << rng.randrange(8)
            try:
                want = po.verify_proof(BB, bytes(bad), inp["program"])
            except po.OracleError as oe:
                with pytest.raises(zlib.ZigzError) as e:
                    zlib.verify_proof(bytes(bad), inp["program"])
                assert e.value.name == oe.name, pos
            else:
                assert zlib.verify_proof(bytes(bad), inp["program"]) == want, pos
    assert zlib.lib().zh_sha256 is not None
    out = np.zeros(32, np.uint8)
    zlib.lib().zh_sha256(b"abc", 3, out.ctypes.data_as(zlib.api.P8))
    assert out.tobytes() == hashlib.sha256(b"abc").digest()


def test_host_pool_and_narrowing_cpp():
    """tests/cpp/test_hostpool.cpp: the upload path's fork-join pool (spin-then-sleep) and the AVX2 u64 -> u32 narrowing
    with its canonical check, compiled and run on the host."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "tests", "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "test_hostpool")
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-pthread", os.path.join(root, "tests", "cpp", "test_hostpool.cpp"),
                    os.path.join(root, "zigz_b200", "csrc", "hostpack.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "hostpool ok" in r.stdout, r.stdout + r.stderr


def test_prodcheck_finish_small_is_a_whole_small_prove(zlib, po):
    """zh_prodcheck_finish_small (the rounds the provers run on the host once their tables are small, four lanes wide from 8
    entries up) from round 0 with a fresh transcript IS a product prove of the small tables: bit-equal to the oracle for every
    degree and size, including edge values (0, p - 1) that exercise the conditional corrections."""
    import ctypes as C
    L = zlib.lib()
    rng = np.random.default_rng(11)
    for d in (1, 2, 3):
        for lg in range(1, 13):  # (a 1-entry table has no rounds: error.NoVariables, sumcheck_prover.zig:30-32)
            m = 1 << lg
            for trial in range(3 if lg < 8 else 1):
                if trial == 0:
                    es = [po.fill_synthetic(BB, 50 + 3 * k + lg, 0, m) for k in range(d)]
                elif trial == 1:
                    es = [rng.choice(np.array([0, 1, BB - 1, BB - 2, (BB + 1) // 2], dtype=np.uint64), size=m) for _ in range(d)]
                else:
                    es = [rng.integers(0, BB, size=m, dtype=np.uint64) for _ in range(d)]
                want = po.prodcheck_prove(BB, es)
                tables = np.ascontiguousarray(np.concatenate(es).astype(np.uint32))
                rp = np.zeros((max(lg, 1), d + 1), np.uint64)
                fp = np.zeros(max(lg, 1), np.uint64)
                fe = np.zeros(3, np.uint64)
                t = zlib.FiatShamirTranscript()
                rc = L.zh_prodcheck_finish_small(d, tables.ctypes.data_as(C.POINTER(C.c_uint32)), m, 0, t._t, None,
                                                 rp.ctypes.data_as(C.POINTER(C.c_uint64)), fp.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                 fe.ctypes.data_as(C.POINTER(C.c_uint64)))
                assert rc == 0
                assert rp[:lg].tolist() == want.round_polys.tolist(), (d, lg, trial)
                assert fp[:lg].tolist() == want.final_point.tolist()
                assert tuple(int(x) for x in fe[:d]) == want.final_evals
    bad = np.array([BB, 0], dtype=np.uint32)
    out = np.zeros(8, np.uint64)
    p64 = out.ctypes.data_as(C.POINTER(C.c_uint64))
    t = zlib.FiatShamirTranscript()
    assert L.zh_prodcheck_finish_small(1, bad.ctypes.data_as(C.POINTER(C.c_uint32)), 2, 0, t._t, None, p64, p64, p64) == -20
    assert L.zh_prodcheck_finish_small(1, bad.ctypes.data_as(C.POINTER(C.c_uint32)), 3, 0, t._t, None, p64, p64, p64) == -22
    assert L.zh_prodcheck_finish_small(4, bad.ctypes.data_as(C.POINTER(C.c_uint32)), 2, 0, t._t, None, p64, p64, p64) == -22


def test_prodcheck_finish_small_vs_golden(zlib, golden):
    """The same host rounds against the independent pure-Python restatement (tests/golden/make_golden.py)."""
    import ctypes as C

    from _cases import synthetic
    L = zlib.lib()
    seen = 0
    for name, case in golden["prodcheck"].items():
        d, n = case["d"], case["n"]
        if n > 4096 or n < 2:
            continue
        lg = n.bit_length() - 1
        tables = np.ascontiguousarray(np.concatenate([synthetic(case["seed"] + k, n) for k in range(d)]).astype(np.uint32))
        rp, fp, fe = np.zeros((lg, d + 1), np.uint64), np.zeros(lg, np.uint64), np.zeros(3, np.uint64)
        t = zlib.FiatShamirTranscript()
        p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
        assert L.zh_prodcheck_finish_small(d, tables.ctypes.data_as(C.POINTER(C.c_uint32)), n, 0, t._t, None, p64(rp), p64(fp), p64(fe)) == 0
        assert rp.tolist() == [list(r) for r in case["round_polys"]], name
        assert fp.tolist() == list(case["final_point"]), name
        seen += 1
    assert seen >= 9
