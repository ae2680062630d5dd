"""LassoProver on the GPU (XXH3 row hashing + sumcheck on the device, flat SHA3 commitments on the host)."""
import numpy as np
import pytest

from _cases import BB, assert_sumcheck_equal, lasso_queries

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("rounds_mode")]


def _same(pr, want):
    assert pr.sumcheck_proof.num_vars == want.sumcheck.num_vars
    assert pr.sumcheck_proof.round_polynomials.tolist() == want.sumcheck.round_polys.tolist()
    assert pr.sumcheck_proof.final_point.tolist() == want.sumcheck.final_point.tolist()
    assert pr.sumcheck_proof.final_eval == want.sumcheck.final_eval
    assert pr.query_commitment == want.query_commitment and pr.table_commitment == want.table_commitment
    assert pr.num_lookups == want.num_lookups


def test_golden(zlib, ctx, golden):
    g = golden["lasso"]
    xor2 = zlib.build_xor_table(2)
    c = g["xor2_queries"]
    for pr in (zlib.LassoProver.prove(ctx, xor2, c["queries"]),
               zlib.LassoProver.prove_with_mapping(ctx, xor2, c["queries"], c["mapping"])):
        assert_sumcheck_equal(pr.sumcheck_proof, c["sumcheck"])
        assert (pr.query_commitment.hex(), pr.table_commitment.hex()) == (c["query_commitment"], c["table_commitment"])
        assert pr.num_lookups == 3
    for op, build, code in (("add", zlib.build_add_table, zlib.TABLE_ADD), ("xor", zlib.build_xor_table, zlib.TABLE_XOR),
                            ("and", zlib.build_and_table, zlib.TABLE_AND)):
        c = g[f"{op}4_200"]
        q = lasso_queries(op, 4, 200)
        for pr in (zlib.LassoProver.prove(ctx, build(4), q), zlib.LassoProver.prove_builtin(ctx, code, 4, q)):
            assert pr.sumcheck_proof.num_vars == 8
            assert_sumcheck_equal(pr.sumcheck_proof, c["sumcheck"])
            assert (pr.query_commitment.hex(), pr.table_commitment.hex()) == (c["query_commitment"], c["table_commitment"])


def test_hash_rows_vs_oracle(zlib, ctx, po):
    rng = np.random.default_rng(0)
    L = zlib.lib()
    import ctypes as C
    for arity in (1, 2, 3, 5):
        rows = rng.integers(0, BB, size=(1000, arity), dtype=np.uint64)
        rows[0] = 0
        rows[1] = BB - 1
        h = C.c_uint64(0)
        ctx.check(L.zb_xxh3_rows(ctx.handle, rows.ctypes.data_as(C.POINTER(C.c_uint64)), 1000, arity, 1024, C.byref(h)))
        got = zlib.Multilinear(ctx, h.value).evaluations
        want = [po.lasso_hash_row(BB, r) for r in rows] + [0] * 24  # zero padding lasso_prover.zig:140-142
        assert got.tolist() == want
    assert zlib.Multilinear(ctx, _table(zlib, ctx, 1, 4)).evaluations.tolist() == \
        [po.lasso_hash_row(BB, r) for r in po.build_table(BB, po.TABLE_XOR, 4)]


def _table(zlib, ctx, op, bits):
    import ctypes as C
    h = C.c_uint64(0)
    ctx.check(zlib.lib().zb_table_mle(ctx.handle, op, bits, C.byref(h)))
    return h.value


@pytest.mark.parametrize("op", ["add", "xor", "and"])
@pytest.mark.parametrize("nq", [2, 3, 1000, 4096, 3 * 1024])
def test_prove_vs_oracle_8bit_tables(zlib, ctx, po, op, nq):
    """RV64I ADD/AND/XOR 8-bit subtables (65536 entries, the XOR8 shape of table_decomposition.zig:130-164)."""
    code = {"add": po.TABLE_ADD, "xor": po.TABLE_XOR, "and": po.TABLE_AND}[op]
    table = po.build_table(BB, code, 8)
    q = lasso_queries(op, 8, nq)
    want = po.lasso_prove(BB, table, q)
    _same(zlib.LassoProver.prove(ctx, table, q), want)
    _same(zlib.LassoProver.prove_builtin(ctx, code, 8, q), want)
    mapping = q[:, 0] * 256 + q[:, 1]
    _same(zlib.LassoProver.prove_with_mapping(ctx, table, q, mapping), want)


def test_errors(zlib, ctx):  # lasso_prover.zig:108-110, 186-201, 352-412
    xor2 = zlib.build_xor_table(2)
    with pytest.raises(zlib.ZigzError) as e:
        zlib.LassoProver.prove(ctx, xor2, np.zeros((0, 3), np.uint64))
    assert e.value.name == "NoQueries"
    with pytest.raises(zlib.ZigzError) as e:  # one query pads to 2^0: SumcheckProver.prove -> NoVariables
        zlib.LassoProver.prove_with_mapping(ctx, xor2, [[3, 2, 1]], [14])
    assert e.value.name == "NoVariables"
    q = [[3, 2, 1], [0, 0, 0]]
    assert zlib.LassoProver.prove_with_mapping(ctx, xor2, q, [14, 0]).num_lookups == 2
    for mapping, name in (([13, 0], "QueryTableMismatch"), ([16, 0], "InvalidMapping"), ([14], "MappingLengthMismatch")):
        with pytest.raises(zlib.ZigzError) as e:
            zlib.LassoProver.prove_with_mapping(ctx, xor2, q, mapping)
        assert e.value.name == name
    with pytest.raises(zlib.ZigzError) as e:  # Multilinear.init(table_evals): 15 entries
        zlib.LassoProver.prove(ctx, xor2[:15], q)
    assert e.value.name == "LengthNotPowerOfTwo"


def test_full_size_lookup_properties(zlib, ctx, po):
    """2^22 lookups (BASELINE config 2) through structure (the bit-for-bit oracle comparison at this size lives in
    test_gpu_baseline_sizes.py): the proof of 2^22 queries that repeat a 2^12 block must (a) pass the verifier's round
    checks and (b) have claimed sum = 2^10 * sum of the block's hashes."""
    blk = lasso_queries("xor", 8, 1 << 12)
    q = np.tile(blk, (1 << 10, 1))
    pr = zlib.LassoProver.prove_builtin(ctx, zlib.TABLE_XOR, 8, q)
    assert pr.sumcheck_proof.num_vars == 22
    block_sum = sum(po.lasso_hash_row(BB, r) for r in blk) % BB
    rp = pr.sumcheck_proof.round_polynomials
    claimed = (2 * int(rp[0][0]) + int(rp[0][1])) % BB
    assert claimed == block_sum * (1 << 10) % BB
    ok, final_claim = po.sumcheck_verify_rounds(BB, rp, claimed)
    assert ok and final_claim == pr.sumcheck_proof.final_eval
    want_small = po.lasso_prove(BB, po.build_table(BB, po.TABLE_XOR, 8), blk)
    assert pr.table_commitment == want_small.table_commitment


@pytest.mark.parametrize("nq,chunk_log2", [(3 * 1024, 8), (4096, 10), (1000, 4), (5, 4)])
def test_pipelined_commitment_matches_oracle(zlib, ctx, po, nq, chunk_log2):
    """The two-thread schedule (query sponge fed by zb_xxh3_rows_stream) forced on small inputs with tiny chunks: same
    proof as the oracle's sequential lasso_prover.zig:103-173, including chunks that end inside a sponge block."""
    table = po.build_table(BB, po.TABLE_XOR, 8)
    q = lasso_queries("xor", 8, nq)
    want = po.lasso_prove(BB, table, q)
    old = zlib.lib().zh_set_lasso_pipeline_min_log2(0)
    ctx.set_option("lasso_chunk_log2", chunk_log2)
    try:
        _same(zlib.LassoProver.prove(ctx, table, q), want)
        _same(zlib.LassoProver.prove_builtin(ctx, po.TABLE_XOR, 8, q), want)
    finally:
        zlib.lib().zh_set_lasso_pipeline_min_log2(old)
        ctx.set_option("lasso_chunk_log2", 19)


def test_pipelined_error_paths(zlib, ctx):
    """A non-canonical row in a late chunk stops the sponge thread and surfaces as the upload's error."""
    old = zlib.lib().zh_set_lasso_pipeline_min_log2(0)
    ctx.set_option("lasso_chunk_log2", 4)
    try:
        q = lasso_queries("xor", 8, 100).copy()
        q[90, 2] = BB  # == p: not canonical
        with pytest.raises(zlib.ZigzError) as e:
            zlib.LassoProver.prove_builtin(ctx, zlib.TABLE_XOR, 8, q)
        assert e.value.name == "NotCanonical"
        ok = zlib.LassoProver.prove_builtin(ctx, zlib.TABLE_XOR, 8, lasso_queries("xor", 8, 100))
        assert ok.num_lookups == 100
    finally:
        zlib.lib().zh_set_lasso_pipeline_min_log2(old)
        ctx.set_option("lasso_chunk_log2", 19)


def test_full_size_commitments_vs_hashlib(zlib, ctx):
    """2^22-entry query polynomial (3 * 2^20 lookups, zero padded): both schedules give the digest hashlib computes over
    the le64 encodings of the downloaded evaluations (commitToPolynomial, lasso_prover.zig:242-252)."""
    import ctypes as C
    import hashlib
    q = np.tile(lasso_queries("and", 8, 3 << 10), (1 << 10, 1))
    h = C.c_uint64(0)
    ctx.check(zlib.lib().zb_xxh3_rows(ctx.handle, q.reshape(-1).ctypes.data_as(C.POINTER(C.c_uint64)), q.shape[0], 3, 1 << 22,
                                      C.byref(h)))
    poly = zlib.Multilinear(ctx, h.value)
    evals = poly.evaluations
    poly.deinit()
    assert not evals[3 << 20:].any()
    want = hashlib.sha3_256(np.ascontiguousarray(evals, dtype="<u8").tobytes()).digest()
    piped = zlib.LassoProver.prove_builtin(ctx, zlib.TABLE_AND, 8, q)
    old = zlib.lib().zh_set_lasso_pipeline_min_log2(-1)
    try:
        serial = zlib.LassoProver.prove_builtin(ctx, zlib.TABLE_AND, 8, q)
    finally:
        zlib.lib().zh_set_lasso_pipeline_min_log2(old)
    assert piped.query_commitment == want and serial.query_commitment == want
    assert piped.table_commitment == serial.table_commitment
    assert piped.sumcheck_proof.to_bytes() == serial.sumcheck_proof.to_bytes()


def test_batch_matches_single_proofs(zlib, ctx, po):
    """zh_lasso_prove_builtin_batch: three tables, ragged query counts, both schedules in one batch — the proofs are
    those of the one-at-a-time calls (and therefore of the oracle, see above)."""
    jobs = [(po.TABLE_ADD, 8, lasso_queries("add", 8, 3000)), (po.TABLE_XOR, 8, lasso_queries("xor", 8, 17)),
            (po.TABLE_AND, 8, lasso_queries("and", 8, 4096)), (po.TABLE_XOR, 4, lasso_queries("xor", 4, 700))]
    singles = [zlib.LassoProver.prove_builtin(ctx, *j) for j in jobs]
    old = zlib.lib().zh_set_lasso_pipeline_min_log2(10)  # jobs 0, 2 and 3 pipelined, job 1 sequential
    ctx.set_option("lasso_chunk_log2", 9)
    try:
        batch = zlib.LassoProver.prove_builtin_batch(ctx, jobs)
        bad = list(jobs)
        q = jobs[2][2].copy()
        q[4000, 0] = BB
        bad[2] = (po.TABLE_AND, 8, q)
        with pytest.raises(zlib.ZigzError) as e:
            zlib.LassoProver.prove_builtin_batch(ctx, bad)
        assert e.value.name == "NotCanonical"
    finally:
        zlib.lib().zh_set_lasso_pipeline_min_log2(old)
        ctx.set_option("lasso_chunk_log2", 19)
    for got, want in zip(batch, singles):
        assert got.sumcheck_proof.to_bytes() == want.sumcheck_proof.to_bytes()
        assert got.query_commitment == want.query_commitment and got.table_commitment == want.table_commitment
        assert got.num_lookups == want.num_lookups
    assert zlib.LassoProver.prove_builtin_batch(ctx, []) == []
