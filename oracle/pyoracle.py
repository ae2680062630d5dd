"""ctypes view of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (zigz_b200/) must not.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzigz_oracle.so")

BABYBEAR_P = 2013265921
F17_P = 17
GOLDILOCKS_P = 0xFFFFFFFF00000001

ERRORS = {
    -1: "EmptyEvaluations", -2: "LengthNotPowerOfTwo", -3: "WrongNumberOfVariables", -4: "NoVariables",
    -5: "EmptyValues", -6: "IndexOutOfBounds", -7: "PointDimensionMismatch", -8: "NoQueries",
    -9: "MappingLengthMismatch", -10: "InvalidMapping", -11: "QueryTableMismatch", -12: "WrongNumberOfChallenges",
    -14: "EmptyTrace", -15: "NoSpaceLeft", -16: "ProgramHashMismatch", -17: "InvalidProof", -100: "OutOfMemory",
}


class OracleError(Exception):
    def __init__(self, code):
        self.code = code
        self.name = ERRORS.get(code, str(code))
        super().__init__(f"error.{self.name}")


def _host_tag() -> str:
    """Identifies the CPU the library was compiled for (-march=native): model name + feature flags."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        model = next((ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith("model name")), "?")
        flags = next((ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith("flags")), "")
        return model + " " + hashlib.sha1(flags.encode()).hexdigest()[:12]
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("zo_hash.c", "zigz_oracle.c", "zo_hash.h", "zigz_oracle.h", "Makefile")]
    tag_file = os.path.join(_HERE, "_build", "host.txt")
    tag = _host_tag()
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    try:
        other_host = open(tag_file).read() != tag  # built with -march=native on another CPU (snapshot from the build container)
    except OSError:
        other_host = True
    if force or stale or other_host:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
        with open(tag_file, "w") as f:
            f.write(tag)
    return _SO


_lib = None
u64 = C.c_uint64
u32 = C.c_uint32
P64 = C.POINTER(C.c_uint64)
P8 = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        L = _lib
        for name in ("zo_f_init", "zo_f_add", "zo_f_sub", "zo_f_mul", "zo_f_pow"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64, u64] if name == "zo_f_init" else [u64, u64, u64]
        L.zo_f_neg.restype = u64
        L.zo_f_neg.argtypes = [u64, u64]
        L.zo_f_inv.argtypes = [u64, u64, P64]
        L.zo_mle_sum.restype = u64
        L.zo_mle_sum.argtypes = [u64, P64, u64]
        L.zo_mle_round_poly.argtypes = [u64, P64, u64, P64]
        L.zo_mle_partial_eval.argtypes = [u64, P64, u64, u64, P64]
        L.zo_mle_eval.argtypes = [u64, P64, u64, P64, u32, P64]
        L.zo_mle_check.argtypes = [u64, C.POINTER(u32)]
        L.zo_transcript_challenge.restype = u64
        L.zo_eval_univariate.restype = u64
        L.zo_eval_univariate.argtypes = [u64, P64, u32, u64]
        L.zo_sumcheck_prove.argtypes = [u64, P64, u64, P64, P64, P64, P64]
        L.zo_sumcheck_prove_interactive.argtypes = [u64, P64, u64, P64, u32, P64, P64, P64]
        L.zo_sumcheck_proof_to_bytes.argtypes = [u32, P64, P64, u64, P8]
        L.zo_sumcheck_verify_rounds.argtypes = [u64, u32, u32, P64, u64, C.POINTER(C.c_int), P64]
        L.zo_prodcheck_prove.argtypes = [u64, C.POINTER(P64), u32, u64, P64, P64, P64, P64]
        L.zo_prod_round_coeffs.argtypes = [u64, C.POINTER(P64), u32, u64, P64]
        L.zo_ceil_pow2.restype = u64
        L.zo_ceil_pow2.argtypes = [u64]
        L.zo_merkle_build.argtypes = [P64, u64, P8, P8, C.POINTER(u32)]
        L.zo_merkle_open.argtypes = [P64, u64, P8, u64, P8, P8, P64]
        L.zo_merkle_verify.argtypes = [P8, u64, P8, P8, u32]
        L.zo_point_to_index.restype = u64
        L.zo_point_to_index.argtypes = [P64, u32]
        L.zo_commit_open.argtypes = [u64, P64, u64, P8, P64, u32, P64, P64, P64, P8, P8]
        L.zo_lasso_hash_row.restype = u64
        L.zo_lasso_hash_row.argtypes = [u64, P64, u32]
        L.zo_lasso_commit_poly.argtypes = [P64, u64, P8]
        L.zo_build_table.argtypes = [u64, C.c_int, u32, P64]
        L.zo_lasso_prove.argtypes = [u64, P64, u64, P64, u64, u32, P64, P64, P64, C.POINTER(u32), P8, P8]
        L.zo_lasso_prove_with_mapping.argtypes = [u64, P64, u64, P64, u64, P64, u64, u32, P64, P64, P64,
                                                  C.POINTER(u32), P8, P8]
        L.zo_generate_commitments.argtypes = [u64, C.c_void_p, C.POINTER(P64), u32, u64, P8, P64, P64, P64, P64, P8, P8]
        L.zo_witness_pack.restype = u64
        L.zo_witness_pack.argtypes = [u64, P64, u64, u32, u32, P64]
        L.zo_count_lookups.restype = u64
        L.zo_count_lookups.argtypes = [P64, u64]
        L.zo_proof_exact_size.restype = C.c_size_t
        L.zo_proof_exact_size.argtypes = [u64, u32, u32, u64]
        L.zo_proof_estimated_size.restype = C.c_size_t
        L.zo_proof_estimated_size.argtypes = [u64, u32, u64]
        L.zo_prove_from_trace.argtypes = [u64, C.c_char_p, C.c_size_t, u64, P64, u32, P64, u64, u64, P64, P64, u32, C.c_int, P8,
                                          C.c_size_t, C.POINTER(C.c_size_t)]
        L.zo_verify_proof.argtypes = [u64, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        L.zo_splitmix64.restype = u64
        L.zo_splitmix64.argtypes = [u64]
        L.zo_fill_synthetic.argtypes = [u64, u64, u64, u64, P64]
        L.zo_xxh3_64_small.restype = u64
        L.zo_xxh3_64_small.argtypes = [C.c_char_p, C.c_size_t, u64]
        L.zo_sha3_256_oneshot.argtypes = [C.c_char_p, C.c_size_t, P8]
        L.zo_sha256_oneshot.argtypes = [C.c_char_p, C.c_size_t, P8]
        L.zo_hash_field_element.argtypes = [u64, P8]
        L.zo_merge_hashes.argtypes = [P8, P8, P8]
    return _lib


def _a(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _p(a: np.ndarray):
    return a.ctypes.data_as(P64)


def _p8(a: np.ndarray):
    return a.ctypes.data_as(P8)


def _chk(rc):
    if rc != 0:
        raise OracleError(rc)


# ---------------------------------------------------------------- hashes
def sha3_256(data: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    lib().zo_sha3_256_oneshot(data, len(data), _p8(out))
    return out.tobytes()


def sha256(data: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    lib().zo_sha256_oneshot(data, len(data), _p8(out))
    return out.tobytes()


def xxh3_64(data: bytes, seed: int = 0) -> int:
    return int(lib().zo_xxh3_64_small(data, len(data), seed))


class Transcript:
    """FiatShamirTranscript (hash.zig:255-324)."""

    class _T(C.Structure):
        _fields_ = [("s", C.c_uint64 * 25), ("buf", C.c_uint8 * 136), ("buf_len", C.c_uint32)]

    def __init__(self):
        self._t = Transcript._T()
        lib().zo_transcript_init(C.byref(self._t))

    def append_field(self, v: int):
        lib().zo_transcript_append_field(C.byref(self._t), u64(v))

    def append_bytes(self, b: bytes):
        lib().zo_transcript_append_bytes(C.byref(self._t), b, C.c_size_t(len(b)))

    def challenge(self, p: int) -> int:
        return int(lib().zo_transcript_challenge(C.byref(self._t), u64(p)))


# ---------------------------------------------------------------- multilinear
def mle_check(n: int) -> int:
    v = u32(0)
    _chk(lib().zo_mle_check(n, C.byref(v)))
    return v.value


def mle_sum(p, evals) -> int:
    e = _a(evals)
    return int(lib().zo_mle_sum(p, _p(e), e.size))


def mle_round_poly(p, evals):
    e = _a(evals)
    out = np.zeros(2, np.uint64)
    _chk(lib().zo_mle_round_poly(p, _p(e), e.size, _p(out)))
    return [int(out[0]), int(out[1])]


def mle_partial_eval(p, evals, r) -> np.ndarray:
    e = _a(evals)
    out = np.zeros(max(e.size // 2, 1), np.uint64)
    _chk(lib().zo_mle_partial_eval(p, _p(e), e.size, r, _p(out)))
    return out


def mle_eval(p, evals, point) -> int:
    e, pt = _a(evals), _a(point)
    out = u64(0)
    _chk(lib().zo_mle_eval(p, _p(e), e.size, _p(pt), pt.size, C.byref(out)))
    return out.value


# ---------------------------------------------------------------- sumcheck
@dataclass
class SumcheckProof:
    num_vars: int
    round_polys: np.ndarray  # (v, ncoef) uint64
    final_point: np.ndarray  # (v,)
    final_eval: int
    claimed_sum: int = 0
    final_evals: tuple = ()

    def to_bytes(self) -> bytes:
        out = np.zeros((2 + 3 * self.num_vars) * 8, np.uint8)
        rp, fp = _a(self.round_polys).reshape(-1), _a(self.final_point)
        lib().zo_sumcheck_proof_to_bytes(self.num_vars, _p(rp), _p(fp), self.final_eval, _p8(out))
        return out.tobytes()


def sumcheck_prove(p, evals) -> SumcheckProof:
    e = _a(evals)
    v = mle_check(e.size)
    rp, fp = np.zeros((max(v, 1), 2), np.uint64), np.zeros(max(v, 1), np.uint64)
    fe, cs = u64(0), u64(0)
    _chk(lib().zo_sumcheck_prove(p, _p(e), e.size, _p(rp), _p(fp), C.byref(fe), C.byref(cs)))
    return SumcheckProof(v, rp[:v], fp[:v], fe.value, cs.value)


def sumcheck_prove_interactive(p, evals, challenges) -> SumcheckProof:
    e, ch = _a(evals), _a(challenges)
    v = mle_check(e.size)
    rp, fp = np.zeros((max(v, 1), 2), np.uint64), np.zeros(max(v, 1), np.uint64)
    fe = u64(0)
    _chk(lib().zo_sumcheck_prove_interactive(p, _p(e), e.size, _p(ch), ch.size, _p(rp), _p(fp), C.byref(fe)))
    return SumcheckProof(v, rp[:v], fp[:v], fe.value)


def sumcheck_verify_rounds(p, round_polys, claimed_sum):
    rp = _a(round_polys)
    v, nc = rp.shape
    ok, fc = C.c_int(0), u64(0)
    _chk(lib().zo_sumcheck_verify_rounds(p, v, nc, _p(rp.reshape(-1)), claimed_sum, C.byref(ok), C.byref(fc)))
    return bool(ok.value), fc.value


def prodcheck_prove(p, polys) -> SumcheckProof:
    arrs = [_a(x) for x in polys]
    d, n = len(arrs), arrs[0].size
    v = mle_check(n)
    ptrs = (P64 * d)(*[_p(a) for a in arrs])
    rp, fp = np.zeros((max(v, 1), d + 1), np.uint64), np.zeros(max(v, 1), np.uint64)
    fes = np.zeros(d, np.uint64)
    cs = u64(0)
    _chk(lib().zo_prodcheck_prove(p, ptrs, d, n, _p(rp), _p(fp), _p(fes), C.byref(cs)))
    fe = 1
    for x in fes:
        fe = fe * int(x) % p
    return SumcheckProof(v, rp[:v], fp[:v], fe, cs.value, tuple(int(x) for x in fes))


def prod_round_coeffs(p, polys) -> list:
    arrs = [_a(x) for x in polys]
    d = len(arrs)
    ptrs = (P64 * d)(*[_p(a) for a in arrs])
    out = np.zeros(d + 1, np.uint64)
    _chk(lib().zo_prod_round_coeffs(p, ptrs, d, arrs[0].size, _p(out)))
    return [int(x) for x in out]


def eval_univariate(p, coeffs, x) -> int:
    c = _a(coeffs)
    return int(lib().zo_eval_univariate(p, _p(c), c.size, x))


# ---------------------------------------------------------------- merkle / commitment
@dataclass
class MerkleTree:
    values: np.ndarray
    leaf_hashes: np.ndarray  # (padded, 32) uint8
    root: bytes
    height: int


def merkle_build(values) -> MerkleTree:
    vals = _a(values)
    if vals.size == 0:
        raise OracleError(-5)
    padded = int(lib().zo_ceil_pow2(vals.size))
    lh = np.zeros((padded, 32), np.uint8)
    root = np.zeros(32, np.uint8)
    h = u32(0)
    _chk(lib().zo_merkle_build(_p(vals), vals.size, _p8(lh), _p8(root), C.byref(h)))
    return MerkleTree(vals, lh, root.tobytes(), h.value)


def merkle_open(tree: MerkleTree, index: int):
    sib = np.zeros((max(tree.height, 1), 32), np.uint8)
    dirs = np.zeros(max(tree.height, 1), np.uint8)
    val = u64(0)
    _chk(lib().zo_merkle_open(_p(tree.values), tree.values.size, _p8(tree.leaf_hashes), index, _p8(sib), _p8(dirs),
                              C.byref(val)))
    return val.value, sib[:tree.height].copy(), dirs[:tree.height].copy()


def merkle_open_many(tree: MerkleTree, indices):
    """[(value, siblings, dirs)] for every index, with one recomputation of the levels."""
    idx = _a(indices)
    k, h = idx.size, max(tree.height, 1)
    sib = np.zeros((k, h, 32), np.uint8)
    dirs = np.zeros((k, h), np.uint8)
    vals = np.zeros(k, np.uint64)
    L = lib()
    L.zo_merkle_open_many.argtypes = [P64, u64, P8, P64, u32, P8, P8, P64]
    _chk(L.zo_merkle_open_many(_p(tree.values), tree.values.size, _p8(tree.leaf_hashes), _p(idx), k, _p8(sib), _p8(dirs), _p(vals)))
    return [(int(vals[j]), sib[j, :tree.height].copy(), dirs[j, :tree.height].copy()) for j in range(k)]


def merkle_verify(root: bytes, value: int, siblings, dirs) -> bool:
    r = np.frombuffer(root, np.uint8).copy()
    s = np.ascontiguousarray(np.asarray(siblings, np.uint8)).reshape(-1)
    d = np.ascontiguousarray(np.asarray(dirs, np.uint8))
    if s.size == 0:
        s, d = np.zeros(32, np.uint8), np.zeros(1, np.uint8)
        h = 0
    else:
        h = d.size
    return bool(lib().zo_merkle_verify(_p8(r), value, _p8(s), _p8(d), h))


def point_to_index(point) -> int:
    pt = _a(point)
    return int(lib().zo_point_to_index(_p(pt), pt.size))


def commit_open(p, tree: MerkleTree, point):
    pt = _a(point)
    sib = np.zeros((max(tree.height, 1), 32), np.uint8)
    dirs = np.zeros(max(tree.height, 1), np.uint8)
    value, li, lv = u64(0), u64(0), u64(0)
    _chk(lib().zo_commit_open(p, _p(tree.values), tree.values.size, _p8(tree.leaf_hashes), _p(pt), pt.size, C.byref(value),
                              C.byref(li), C.byref(lv), _p8(sib), _p8(dirs)))
    return value.value, li.value, lv.value, sib[:tree.height].copy(), dirs[:tree.height].copy()


def hash_leaf(value: int) -> bytes:
    out = np.zeros(32, np.uint8)
    lib().zo_hash_field_element(value, _p8(out))
    return out.tobytes()


@dataclass
class Commitments:
    roots: np.ndarray         # (count, 32)
    points: np.ndarray        # (count, v)
    values: np.ndarray        # (count,)
    leaf_indices: np.ndarray  # (count,)
    leaf_values: np.ndarray   # (count,)
    siblings: np.ndarray      # (count, v, 32)
    dirs: np.ndarray          # (count, v)


def generate_commitments(p, transcript: "Transcript", polys) -> Commitments:
    """Prover.generateCommitments (prover.zig:366-467); the transcript is advanced exactly as the reference does."""
    arrs = [_a(x) for x in polys]
    k, n = len(arrs), arrs[0].size
    v = mle_check(n)
    ptrs = (P64 * k)(*[_p(a) for a in arrs])
    vv = max(v, 1)
    out = Commitments(np.zeros((k, 32), np.uint8), np.zeros((k, vv), np.uint64), np.zeros(k, np.uint64), np.zeros(k, np.uint64),
                      np.zeros(k, np.uint64), np.zeros((k, vv, 32), np.uint8), np.zeros((k, vv), np.uint8))
    # the C side strides by v (not max(v,1)): use flat buffers of the exact size
    pts, sib, dirs = np.zeros(k * vv, np.uint64), np.zeros(k * vv * 32, np.uint8), np.zeros(k * vv, np.uint8)
    _chk(lib().zo_generate_commitments(p, C.byref(transcript._t), ptrs, k, n, _p8(out.roots), _p(pts), _p(out.values),
                                       _p(out.leaf_indices), _p(out.leaf_values), _p8(sib), _p8(dirs)))
    out.points = pts[:k * v].reshape(k, v)
    out.siblings = sib[:k * v * 32].reshape(k, v, 32)
    out.dirs = dirs[:k * v].reshape(k, v)
    return out


def witness_pack(p, cols, n_hold=33) -> np.ndarray:
    """cols: (n_cols, num_steps) raw u64 -> (n_cols, padded) canonical evaluations (witness.zig:29-270)."""
    c = _a(cols)
    n_cols, steps = c.shape
    padded = 1
    while padded < steps:
        padded <<= 1
    out = np.zeros((n_cols, padded), np.uint64)
    got = lib().zo_witness_pack(p, _p(c.reshape(-1)) if c.size else _p(np.zeros(1, np.uint64)), steps, n_cols, n_hold, _p(out.reshape(-1)))
    assert got == padded
    return out


VERDICTS = {0: "Accept", 1: "RejectInvalidSumcheck", 2: "RejectInvalidLookup", 3: "RejectInvalidCommitment"}


def prove_from_trace(p, program: bytes, entry_pc, initial_regs, cols, final_pc, final_regs, outputs, compat_buffer=False) -> bytes:
    """Prover.prove after the VM + BinarySerializer.serialize (prover.zig:91-226, serialization.zig:70-97)."""
    c = _a(cols)
    steps = c.shape[1] if c.ndim == 2 else 0
    ir, fr, out = _a(initial_regs if initial_regs is not None else []), _a(final_regs), _a(outputs if outputs is not None else [])
    n_lookups = int(lib().zo_count_lookups(_p(np.ascontiguousarray(c[33])), steps)) if steps else 0
    cap = int(lib().zo_proof_exact_size(max(steps, 1), ir.size, out.size, n_lookups))
    buf = np.zeros(cap, np.uint8)
    n = C.c_size_t(0)
    z1 = np.zeros(1, np.uint64)
    _chk(lib().zo_prove_from_trace(p, program, len(program), entry_pc, _p(ir) if ir.size else _p(z1), ir.size,
                                   _p(c.reshape(-1)) if c.size else _p(z1), steps, final_pc, _p(fr), _p(out) if out.size else _p(z1),
                                   out.size, 1 if compat_buffer else 0, _p8(buf), cap, C.byref(n)))
    return buf[:n.value].tobytes()


def verify_proof(p, proof: bytes, program: bytes) -> str:
    res = C.c_int(-1)
    _chk(lib().zo_verify_proof(p, proof, len(proof), program, len(program), C.byref(res)))
    return VERDICTS[res.value]


# ---------------------------------------------------------------- lasso
TABLE_ADD, TABLE_XOR, TABLE_AND = 0, 1, 2


def build_table(p, op, bits) -> np.ndarray:
    rows = np.zeros((1 << (2 * bits), 3), np.uint64)
    lib().zo_build_table(p, op, bits, _p(rows))
    return rows


def lasso_hash_row(p, row) -> int:
    r = _a(row)
    return int(lib().zo_lasso_hash_row(p, _p(r), r.size))


def lasso_commit_poly(evals) -> bytes:
    e = _a(evals)
    out = np.zeros(32, np.uint8)
    lib().zo_lasso_commit_poly(_p(e), e.size, _p8(out))
    return out.tobytes()


@dataclass
class LassoProof:
    sumcheck: SumcheckProof
    query_commitment: bytes
    table_commitment: bytes
    num_lookups: int


def lasso_prove(p, table_rows, query_rows, mapping=None) -> LassoProof:
    t, q = _a(table_rows), _a(query_rows)
    arity = t.shape[1] if t.ndim == 2 else q.shape[1]
    nq = q.shape[0] if q.ndim == 2 else 0
    nt = t.shape[0]
    vmax = max(int(lib().zo_ceil_pow2(max(nq, 1))).bit_length(), 1)
    rp, fp = np.zeros((vmax, 2), np.uint64), np.zeros(vmax, np.uint64)
    fe, nv = u64(0), u32(0)
    qc, tc = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
    tq = q.reshape(-1) if q.size else np.zeros(1, np.uint64)
    if mapping is None:
        rc = lib().zo_lasso_prove(p, _p(t.reshape(-1)), nt, _p(tq), nq, arity, _p(rp), _p(fp), C.byref(fe), C.byref(nv),
                                  _p8(qc), _p8(tc))
    else:
        m = _a(mapping)
        mm = m if m.size else np.zeros(1, np.uint64)
        rc = lib().zo_lasso_prove_with_mapping(p, _p(t.reshape(-1)), nt, _p(tq), nq, _p(mm), m.size, arity, _p(rp), _p(fp),
                                               C.byref(fe), C.byref(nv), _p8(qc), _p8(tc))
    _chk(rc)
    v = nv.value
    return LassoProof(SumcheckProof(v, rp[:v], fp[:v], fe.value), qc.tobytes(), tc.tobytes(), nq)


# ---------------------------------------------------------------- synthetic inputs
def fill_synthetic(p, seed, start, n) -> np.ndarray:
    out = np.zeros(n, np.uint64)
    lib().zo_fill_synthetic(p, seed, start, n, _p(out))
    return out


def splitmix64(x: int) -> int:
    return int(lib().zo_splitmix64(x & 0xFFFFFFFFFFFFFFFF))
