"""Python view of the zigz prover-side API, names and semantics as in the reference's src/lib.zig:9-17.

Every class below is a thin ctypes shim over the C ABI (include/zigz_b200.h, include/zigz_host.h); all arithmetic
runs either in the sm_100a kernels or in the C++ host twin.  Field elements are Python ints / numpy uint64 holding
canonical BabyBear values, the reference's `struct { value: u64 }` (src/core/field.zig:26-27).
Zig error unions surface as `ZigzError` with the reference's error name (`.name`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from ._cabi import ZigzError, lib

BABYBEAR_P = 2013265921  # src/core/field_presets.zig:19

u64 = C.c_uint64
u32 = C.c_uint32
P64 = C.POINTER(C.c_uint64)
P8 = C.POINTER(C.c_uint8)


def _a64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _p64(a: np.ndarray):
    return a.ctypes.data_as(P64)


def _p8(a: np.ndarray):
    return a.ctypes.data_as(P8)


def _b8(b: bytes):
    """bytes -> (keep-alive array, uint8_t*)"""
    a = np.frombuffer(b if len(b) else b"\x00", np.uint8)
    return a, a.ctypes.data_as(P8)


class Context:
    """One GPU + one stream; calls are synchronous and serialized like the single-threaded reference."""

    def __init__(self, device: int = 0, device_mask: Optional[int] = None):
        """device: one GPU. device_mask (bit d = GPU d, 2/4/8 bits): ONE context over several GPUs of this process
        (zb_ctx_create_mask): tables are sharded, provers and commitments run on all of them, results are unchanged."""
        self._h = C.c_void_p()
        if device_mask is not None:
            rc = lib().zb_ctx_create_mask(device_mask, C.byref(self._h))
        else:
            rc = lib().zb_ctx_create(device, C.byref(self._h))
        if rc != 0:
            self._h = None
            raise ZigzError(rc, "zb_ctx_create: a CUDA device is required, there is no CPU fallback")

    @property
    def n_devices(self) -> int:
        return int(lib().zb_group_size(self._h))

    def check(self, rc: int):
        if rc != 0:
            raise ZigzError(rc, lib().zb_last_error(self._h).decode() if rc <= -200 else "")

    @property
    def handle(self):
        return self._h

    @property
    def kernel_launches(self) -> int:
        return int(lib().zb_kernel_launches(self._h))

    @property
    def stream(self) -> int:
        return int(lib().zb_stream(self._h) or 0)

    def sync(self):
        self.check(lib().zb_sync(self._h))

    def set_option(self, key: str, value: int):
        self.check(lib().zb_set_option(self._h, key.encode(), value))

    def get_option(self, key: str) -> int:
        v = C.c_int64(0)
        self.check(lib().zb_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    # ---- multi-GPU: NCCL communicator attached to this context (one process per GPU)
    def comm_init(self, unique_id: bytes, rank: int, world: int, nccl_path: Optional[str] = None):
        uid = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self.check(lib().zb_comm_init(self._h, nccl_path.encode() if nccl_path else None, uid, rank, world))

    @staticmethod
    def comm_unique_id(nccl_path: Optional[str] = None) -> bytes:
        uid = (C.c_uint8 * 128)()
        rc = lib().zb_comm_unique_id(nccl_path.encode() if nccl_path else None, uid)
        if rc != 0:
            raise ZigzError(rc, "zb_comm_unique_id")
        return bytes(uid)

    def p2p_handle(self) -> bytes:
        h = (C.c_uint8 * 64)()
        self.check(lib().zb_comm_p2p_handle(self._h, h))
        return bytes(h)

    def p2p_attach(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.check(lib().zb_comm_p2p_attach(self._h, buf))

    @property
    def world(self) -> int:
        r, w = C.c_int32(0), C.c_int32(1)
        lib().zb_comm_info(self._h, C.byref(r), C.byref(w))
        return w.value

    @property
    def rank(self) -> int:
        r, w = C.c_int32(0), C.c_int32(1)
        lib().zb_comm_info(self._h, C.byref(r), C.byref(w))
        return r.value

    def allreduce_u64(self, vals) -> np.ndarray:
        a = _a64(vals).copy()
        self.check(lib().zb_comm_allreduce_u64(self._h, _p64(a), a.size))
        return a

    def int_pipe_peak(self):
        """Measured 32-bit lane-ops/s of LOP3 chains, SHF chains and the Keccak mix on this GPU."""
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        self.check(lib().zb_int_pipe_peak(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"lop3_per_s": a.value, "shf_per_s": b.value, "keccak_mix_per_s": c.value}

    def h2d_rate(self, nbytes: int = 1 << 30) -> float:
        """Measured host->device copy rate (bytes/s) from pinned memory on this GPU's link."""
        v = C.c_double(0)
        self.check(lib().zb_h2d_rate(self._h, nbytes, C.byref(v)))
        return v.value

    def device_info(self):
        sm, tot, free = C.c_int32(0), u64(0), u64(0)
        self.check(lib().zb_device_info(self._h, C.byref(sm), C.byref(tot), C.byref(free)))
        return {"sm_count": sm.value, "total_mem": tot.value, "free_mem": free.value}

    def pinned(self, n: int, dtype=np.uint64) -> np.ndarray:
        """numpy array backed by page-locked host memory (freed with the context)."""
        nbytes = int(n) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(lib().zb_host_alloc(self._h, max(nbytes, 1), C.byref(p)))
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return arr

    def timer_start(self):
        self.check(lib().zb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self.check(lib().zb_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def profile(self, on: bool):
        self.check(lib().zb_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{kernel: (launches, total_ms, algorithmic_bytes)} accumulated since profile(True)."""
        n = lib().zb_profile_count(self._h)
        out = {}
        for i in range(n):
            name = C.create_string_buffer(64)
            cnt, ms, by = u64(0), C.c_double(0), u64(0)
            self.check(lib().zb_profile_entry(self._h, i, name, 64, C.byref(cnt), C.byref(ms), C.byref(by)))
            out[name.value.decode()] = (cnt.value, ms.value, by.value)
        return out

    def close(self):
        if self._h:
            for p in getattr(self, "_pinned", []):
                lib().zb_host_free(self._h, p)
            self._pinned = []
            lib().zb_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------- FiatShamirTranscript
class FiatShamirTranscript:
    """src/core/hash.zig:255-324 — host-resident."""

    def __init__(self):
        self._t = lib().zh_transcript_new()

    def append_field_element(self, v: int):
        lib().zh_transcript_append_field(self._t, v)

    def append_field_elements(self, vs):
        a = _a64(vs)
        lib().zh_transcript_append_fields(self._t, _p64(a), a.size)

    def append_bytes(self, b: bytes):
        lib().zh_transcript_append_bytes(self._t, b, len(b))

    def challenge(self) -> int:
        return int(lib().zh_transcript_challenge(self._t))

    def finalize(self) -> bytes:
        out = np.zeros(32, np.uint8)
        lib().zh_transcript_finalize(self._t, _p8(out))
        return out.tobytes()

    def __del__(self):
        if getattr(self, "_t", None):
            lib().zh_transcript_free(self._t)
            self._t = None


def sha3_256(data: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    lib().zh_sha3_256(data, len(data), _p8(out))
    return out.tobytes()


def eval_univariate_coeffs(coeffs, x: int) -> int:
    """src/proofs/sumcheck_protocol.zig:113-123"""
    c = _a64(coeffs)
    return int(lib().zh_eval_univariate(_p64(c), c.size, x))


# --------------------------------------------------------------------------- Multilinear
class Multilinear:
    """src/poly/multilinear.zig:20 — dense MLE resident in HBM as canonical u32."""

    def __init__(self, ctx: Context, handle: int):
        self.ctx, self._h = ctx, handle

    # init :36-54
    @classmethod
    def init(cls, ctx: Context, evaluations) -> "Multilinear":
        e = _a64(evaluations)
        h = u64(0)
        ctx.check(lib().zb_mle_upload(ctx.handle, _p64(e), e.size, C.byref(h)))
        return cls(ctx, h.value)

    @classmethod
    def init_u32(cls, ctx: Context, evaluations: np.ndarray) -> "Multilinear":
        e = np.ascontiguousarray(evaluations, dtype=np.uint32)
        h = u64(0)
        ctx.check(lib().zb_mle_upload_u32(ctx.handle, e.ctypes.data_as(C.POINTER(C.c_uint32)), e.size, C.byref(h)))
        return cls(ctx, h.value)

    @classmethod
    def zero(cls, ctx, num_vars):  # :57-70
        return cls.constant(ctx, num_vars, 0)

    @classmethod
    def constant(cls, ctx, num_vars, value):  # :73-86
        h = u64(0)
        ctx.check(lib().zb_mle_constant(ctx.handle, num_vars, value, C.byref(h)))
        return cls(ctx, h.value)

    @classmethod
    def synthetic(cls, ctx, seed, n, start=0, stride=1):
        """e[i] = splitmix64(seed + start + i*stride) mod p generated on the device."""
        h = u64(0)
        ctx.check(lib().zb_mle_synthetic(ctx.handle, seed, start, stride, n, C.byref(h)))
        return cls(ctx, h.value)

    def clone(self) -> "Multilinear":
        h = u64(0)
        self.ctx.check(lib().zb_mle_clone(self.ctx.handle, self._h, C.byref(h)))
        return Multilinear(self.ctx, h.value)

    def deinit(self):  # :89-91
        if self._h:
            lib().zb_mle_free(self.ctx.handle, self._h)
            self._h = 0

    @property
    def handle(self) -> int:
        return self._h

    def _len(self):
        n, v = u64(0), u32(0)
        self.ctx.check(lib().zb_mle_len(self.ctx.handle, self._h, C.byref(n), C.byref(v)))
        return n.value, v.value

    @property
    def num_vars(self) -> int:
        return self._len()[1]

    def __len__(self):
        return self._len()[0]

    @property
    def evaluations(self) -> np.ndarray:
        n = len(self)
        out = np.zeros(n, np.uint64)
        self.ctx.check(lib().zb_mle_download(self.ctx.handle, self._h, _p64(out), n))
        return out

    def eval(self, point) -> int:  # :110-144 (LSB-first)
        pt = _a64(point)
        out = u64(0)
        self.ctx.check(lib().zb_mle_eval(self.ctx.handle, self._h, _p64(pt) if pt.size else None, pt.size, C.byref(out)))
        return out.value

    def partial_eval(self, r: int, with_next_sums: bool = False):  # :154-180 (binds the top index bit)
        h = u64(0)
        nxt = (u64 * 2)()
        self.ctx.check(lib().zb_mle_partial_eval(self.ctx.handle, self._h, r, C.byref(h), nxt))
        m = Multilinear(self.ctx, h.value)
        return (m, (nxt[0], nxt[1])) if with_next_sums else m

    def fold_inplace(self, r: int):
        nxt = (u64 * 2)()
        self.ctx.check(lib().zb_mle_fold_inplace(self.ctx.handle, self._h, r, nxt))
        return nxt[0], nxt[1]

    def sum_over_hypercube(self) -> int:  # :188-194
        out = u64(0)
        self.ctx.check(lib().zb_mle_sum(self.ctx.handle, self._h, C.byref(out)))
        return out.value

    def round_polynomial(self) -> List[int]:  # :205-232 -> [s0, s1 - s0]
        s = (u64 * 2)()
        self.ctx.check(lib().zb_mle_round_sums(self.ctx.handle, self._h, s))
        return [int(s[0]), int(lib().zh_f_sub(s[1], s[0]))]

    def add(self, other: "Multilinear") -> "Multilinear":  # :235-250
        h = u64(0)
        self.ctx.check(lib().zb_mle_add(self.ctx.handle, self._h, other._h, C.byref(h)))
        return Multilinear(self.ctx, h.value)

    def scalar_mul(self, s: int) -> "Multilinear":  # :253-264
        h = u64(0)
        self.ctx.check(lib().zb_mle_scalar_mul(self.ctx.handle, self._h, s, C.byref(h)))
        return Multilinear(self.ctx, h.value)


# --------------------------------------------------------------------------- Sumcheck
@dataclass
class SumcheckProof:
    """src/proofs/sumcheck_protocol.zig:24-109"""
    num_vars: int
    round_polynomials: np.ndarray  # (v, ncoef) uint64
    final_point: np.ndarray        # (v,)
    final_eval: int
    claimed_sum: Optional[int] = None
    final_evals: tuple = ()

    def to_bytes(self) -> bytes:  # toBytes :76-109
        out = np.zeros((2 + 3 * self.num_vars) * 8, np.uint8)
        rp, fp = _a64(self.round_polynomials).reshape(-1), _a64(self.final_point)
        n = lib().zh_sumcheck_proof_to_bytes(self.num_vars, _p64(rp), _p64(fp), self.final_eval, _p8(out))
        return out[:n].tobytes()


class SumcheckProver:
    """src/proofs/sumcheck_prover.zig — SumcheckProver(BabyBear)"""

    @staticmethod
    def prove(poly: Multilinear) -> SumcheckProof:  # :26-91
        ctx = poly.ctx
        v = max(poly.num_vars + (ctx.world - 1).bit_length(), 1)
        rp, fp = np.zeros((v, 2), np.uint64), np.zeros(v, np.uint64)
        fe, cs = u64(0), u64(0)
        ctx.check(lib().zh_sumcheck_prove(ctx.handle, poly.handle, _p64(rp), _p64(fp), C.byref(fe), C.byref(cs)))
        return SumcheckProof(v, rp, fp, fe.value, cs.value)

    @staticmethod
    def prove_interactive(poly: Multilinear, challenges) -> SumcheckProof:  # :97-144
        ctx = poly.ctx
        ch = _a64(challenges)
        v = max(poly.num_vars, 1)
        rp, fp = np.zeros((v, 2), np.uint64), np.zeros(v, np.uint64)
        fe = u64(0)
        ctx.check(lib().zh_sumcheck_prove_interactive(ctx.handle, poly.handle, _p64(ch) if ch.size else None, ch.size,
                                                      _p64(rp), _p64(fp), C.byref(fe)))
        return SumcheckProof(v, rp, fp, fe.value)


class EqProductSumcheckProver:
    """EXTENSION (no reference behaviour): eq-weighted product sumcheck sum_x eq(tau, x) prod_k A_k(x), d = 1, 2."""

    @staticmethod
    def prove(tau, polys: Sequence["Multilinear"]) -> "SumcheckProof":
        ctx = polys[0].ctx
        d = len(polys)
        v = max(polys[0].num_vars, 1)
        t = _a64(tau)
        hs = (u64 * d)(*[p.handle for p in polys])
        rp, fp, fes = np.zeros((v, d + 2), np.uint64), np.zeros(v, np.uint64), np.zeros(d + 1, np.uint64)
        cs = u64(0)
        ctx.check(lib().zh_eqcheck_prove(ctx.handle, _p64(t) if t.size else None, t.size, hs, d, _p64(rp), _p64(fp), _p64(fes), C.byref(cs)))
        fe = 1
        for x in fes:
            fe = fe * int(x) % BABYBEAR_P
        return SumcheckProof(v, rp, fp, fe, cs.value, tuple(int(x) for x in fes))


class ProductSumcheckProver:
    """Product of d (1..3) multilinears — extension in the reference's conventions."""

    @staticmethod
    def prove(polys: Sequence[Multilinear], consume: bool = False) -> SumcheckProof:
        ctx = polys[0].ctx
        d = len(polys)
        v = max(polys[0].num_vars + (ctx.world - 1).bit_length(), 1)  # sharded: v_local + log2(world) rounds
        hs = (u64 * d)(*[p.handle for p in polys])
        rp, fp, fes = np.zeros((v, d + 1), np.uint64), np.zeros(v, np.uint64), np.zeros(d, np.uint64)
        cs = u64(0)
        fn = lib().zh_prodcheck_prove_consume if consume else lib().zh_prodcheck_prove
        ctx.check(fn(ctx.handle, hs, d, _p64(rp), _p64(fp), _p64(fes), C.byref(cs)))
        fe = 1
        for x in fes:
            fe = fe * int(x) % BABYBEAR_P
        return SumcheckProof(v, rp, fp, fe, cs.value, tuple(int(x) for x in fes))


# --------------------------------------------------------------------------- Merkle / commitments
@dataclass
class MerklePath:
    """src/commitments/merkle_tree.zig:39-60"""
    siblings: np.ndarray    # (height, 32) uint8, leaf -> root
    directions: np.ndarray  # (height,) uint8: 1 = this node is the right child


@dataclass
class MerkleOpeningProof:
    """src/commitments/merkle_tree.zig:63-75"""
    value: int
    index: int
    path: MerklePath


class SimpleMerkleTree:
    """src/commitments/merkle_tree.zig:273-402 with SHA3Hasher; all levels retained in HBM."""

    def __init__(self, ctx: Context, handle: int, root: bytes):
        self.ctx, self._h, self._root = ctx, handle, root

    @classmethod
    def build(cls, ctx: Context, values) -> "SimpleMerkleTree":  # :283-318
        vals = _a64(values)
        h = u64(0)
        root = np.zeros(32, np.uint8)
        ctx.check(lib().zb_merkle_build_values(ctx.handle, _p64(vals) if vals.size else None, vals.size, C.byref(h), _p8(root)))
        return cls(ctx, h.value, root.tobytes())

    @property
    def handle(self):
        return self._h

    def get_root(self) -> bytes:  # :320-322
        return self._root

    def _info(self):
        n, hgt = u64(0), u32(0)
        self.ctx.check(lib().zb_merkle_info(self.ctx.handle, self._h, C.byref(n), C.byref(hgt), None))
        return n.value, hgt.value

    @property
    def height(self) -> int:
        return self._info()[1]

    def open(self, index: int) -> MerkleOpeningProof:  # :324-360
        _, hgt = self._info()
        sib = np.zeros((max(hgt, 1), 32), np.uint8)
        dirs = np.zeros(max(hgt, 1), np.uint8)
        val = u64(0)
        self.ctx.check(lib().zb_merkle_open(self.ctx.handle, self._h, index, _p8(sib), _p8(dirs), C.byref(val)))
        return MerkleOpeningProof(val.value, index, MerklePath(sib[:hgt].copy(), dirs[:hgt].copy()))

    def leaf_hashes(self) -> np.ndarray:
        n, hgt = self._info()
        padded = 1 << hgt
        out = np.zeros((padded, 32), np.uint8)
        self.ctx.check(lib().zb_merkle_leaf_hashes(self.ctx.handle, self._h, _p8(out), padded))
        return out

    @staticmethod
    def verify(root: bytes, proof: MerkleOpeningProof) -> bool:  # :362-373
        r = np.frombuffer(root, np.uint8).copy()
        h = proof.path.directions.size
        s = np.ascontiguousarray(proof.path.siblings, np.uint8).reshape(-1) if h else np.zeros(32, np.uint8)
        d = np.ascontiguousarray(proof.path.directions, np.uint8) if h else np.zeros(1, np.uint8)
        return bool(lib().zh_merkle_verify(_p8(r), proof.value, _p8(s), _p8(d), h))

    def deinit(self):
        if self._h:
            lib().zb_merkle_free(self.ctx.handle, self._h)
            self._h = 0


class MerkleTree(SimpleMerkleTree):
    """src/commitments/merkle_tree.zig:78-264 — the pointer-based variant. It pads to a power of two with hash(0) and
    splits at the middle recursively (:87-110, :187-215), which yields exactly SimpleMerkleTree's root; its `open` is
    unusable in the reference (`getValueAtIndex` -> error.NotImplemented, :237-243), so only build / root_hash / verify
    are offered, aliased onto the same device tree."""

    @classmethod
    def build(cls, ctx: Context, values) -> "MerkleTree":
        t = SimpleMerkleTree.build(ctx, values)
        return cls(t.ctx, t.handle, t.get_root())

    def root_hash(self) -> bytes:  # :113-115
        return self.get_root()

    def open(self, index: int):  # :118-150 -> getValueAtIndex :237-243
        raise ZigzError(-22, "MerkleTree.open is error.NotImplemented in the reference; use SimpleMerkleTree")


@dataclass
class PolynomialCommitment:
    """src/commitments/polynomial_commit.zig:24-39"""
    commitment: bytes
    num_vars: int


@dataclass
class OpeningProof:
    """src/commitments/polynomial_commit.zig:42-55"""
    point: np.ndarray
    value: int
    merkle_proof: MerkleOpeningProof


class CommitmentScheme:
    """src/commitments/polynomial_commit.zig:58-185 — CommitmentSchemeSHA3(BabyBear)"""

    @staticmethod
    def commit(poly: Multilinear):  # :69-83
        ctx = poly.ctx
        h, v = u64(0), u32(0)
        root = np.zeros(32, np.uint8)
        ctx.check(lib().zh_commit(ctx.handle, poly.handle, C.byref(h), _p8(root), C.byref(v)))
        return PolynomialCommitment(root.tobytes(), v.value), SimpleMerkleTree(ctx, h.value, root.tobytes())

    @staticmethod
    def commit_sharded(local_poly: Multilinear):
        """Contiguous-block shard of a polynomial spread over the context's communicator -> (global commitment, local tree)."""
        ctx = local_poly.ctx
        h = u64(0)
        lroot, root = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
        ctx.check(lib().zh_commit_sharded(ctx.handle, local_poly.handle, C.byref(h), _p8(lroot), _p8(root)))
        v = local_poly.num_vars + (ctx.world - 1).bit_length()
        return PolynomialCommitment(root.tobytes(), v), SimpleMerkleTree(ctx, h.value, lroot.tobytes())

    @staticmethod
    def batch_commit(polys: Sequence[Multilinear]):  # :132-157
        ctx = polys[0].ctx
        k = len(polys)
        hs = (u64 * k)(*[p.handle for p in polys])
        ts = (u64 * k)()
        roots = np.zeros((k, 32), np.uint8)
        ctx.check(lib().zh_batch_commit(ctx.handle, hs, k, ts, _p8(roots)))
        v = polys[0].num_vars
        return ([PolynomialCommitment(roots[i].tobytes(), v) for i in range(k)],
                [SimpleMerkleTree(ctx, ts[i], roots[i].tobytes()) for i in range(k)])

    @staticmethod
    def open(poly: Multilinear, tree: SimpleMerkleTree, point) -> OpeningProof:  # :86-115
        ctx = poly.ctx
        pt = _a64(point)
        hgt = max(tree.height, 1)
        sib, dirs = np.zeros((hgt, 32), np.uint8), np.zeros(hgt, np.uint8)
        val, li, lv = u64(0), u64(0), u64(0)
        ctx.check(lib().zh_commit_open(ctx.handle, poly.handle, tree.handle, _p64(pt) if pt.size else None, pt.size,
                                       C.byref(val), C.byref(li), C.byref(lv), _p8(sib), _p8(dirs)))
        h = tree.height
        return OpeningProof(pt.copy(), val.value, MerkleOpeningProof(lv.value, li.value, MerklePath(sib[:h].copy(), dirs[:h].copy())))

    @staticmethod
    def verify(commitment: PolynomialCommitment, proof: OpeningProof) -> bool:  # :118-129
        if proof.point.size != commitment.num_vars:
            return False
        return SimpleMerkleTree.verify(commitment.commitment, proof.merkle_proof)

    @staticmethod
    def batch_verify(commitments, proofs) -> bool:  # :160-175
        if len(commitments) != len(proofs):
            return False
        return all(CommitmentScheme.verify(c, p) for c, p in zip(commitments, proofs))

    @staticmethod
    def point_to_index(point) -> int:  # :178-183
        pt = _a64(point)
        return int(lib().zh_point_to_index(_p64(pt) if pt.size else None, pt.size))


@dataclass
class CommitmentOpenings:
    """The `witness_commitments` of a proof (src/prover/proof.zig:147-190), struct-of-arrays."""
    roots: np.ndarray         # (count, 32)
    points: np.ndarray        # (count, v)
    values: np.ndarray        # (count,)
    leaf_indices: np.ndarray  # (count,)
    leaf_values: np.ndarray   # (count,)
    siblings: np.ndarray      # (count, v, 32)
    dirs: np.ndarray          # (count, v)


def generate_commitments(transcript: FiatShamirTranscript, polys: Sequence[Multilinear]) -> CommitmentOpenings:
    """Prover.generateCommitments (src/prover/prover.zig:366-467): one batched commit, transcript interleave, openings."""
    ctx = polys[0].ctx
    k, v = len(polys), polys[0].num_vars
    hs = (u64 * k)(*[p.handle for p in polys])
    vv = max(v, 1)
    roots = np.zeros((k, 32), np.uint8)
    pts, vals = np.zeros(k * vv, np.uint64), np.zeros(k, np.uint64)
    li, lv = np.zeros(k, np.uint64), np.zeros(k, np.uint64)
    sib, dirs = np.zeros(k * vv * 32, np.uint8), np.zeros(k * vv, np.uint8)
    ctx.check(lib().zh_generate_commitments(ctx.handle, transcript._t, hs, k, _p8(roots), _p64(pts), _p64(vals), _p64(li), _p64(lv),
                                            _p8(sib), _p8(dirs)))
    return CommitmentOpenings(roots, pts[:k * v].reshape(k, v), vals, li, lv, sib[:k * v * 32].reshape(k, v, 32),
                              dirs[:k * v].reshape(k, v))


WITNESS_COLUMNS = (["pc"] + [f"x{i}" for i in range(32)] +
                   ["opcode", "rd", "rs1", "rs2", "funct3", "funct7", "imm", "mem_address", "mem_value", "mem_is_read"])


def witness_pack(ctx: Context, cols, n_hold: int = 33) -> List[Multilinear]:
    """WitnessGenerator.generate (src/constraints/witness.zig:29-270) from SoA trace columns: cols is (n_cols, num_steps)
    raw u64 in the order WITNESS_COLUMNS (= prover.zig:376-390). Returns the n_cols padded polynomials."""
    c = _a64(cols)
    n_cols, steps = c.shape
    out = (u64 * n_cols)()
    nv = u32(0)
    ctx.check(lib().zb_witness_pack(ctx.handle, _p64(c.reshape(-1)) if c.size else None, steps, n_cols, n_hold, out, C.byref(nv)))
    return [Multilinear(ctx, out[i]) for i in range(n_cols)]


VERDICTS = {0: "Accept", 1: "RejectInvalidSumcheck", 2: "RejectInvalidLookup", 3: "RejectInvalidCommitment"}


def witness_pack_commit(ctx: Context, cols, n_hold: int = 33):
    """witness_pack + CommitmentScheme.batch_commit as one pipeline (zb_witness_pack_commit: the trace upload overlaps the
    leaf hashing). Returns (polynomials, commitments, trees) — same values as the two calls in a row."""
    c = _a64(cols)
    n_cols, steps = c.shape
    out, trees = (u64 * n_cols)(), (u64 * n_cols)()
    roots = np.zeros((n_cols, 32), np.uint8)
    nv = u32(0)
    ctx.check(lib().zb_witness_pack_commit(ctx.handle, _p64(c.reshape(-1)), steps, n_cols, n_hold, out, C.byref(nv), trees, _p8(roots)))
    polys = [Multilinear(ctx, out[i]) for i in range(n_cols)]
    return (polys, [PolynomialCommitment(roots[i].tobytes(), nv.value) for i in range(n_cols)],
            [SimpleMerkleTree(ctx, trees[i], roots[i].tobytes()) for i in range(n_cols)])


def prove_from_trace(ctx: Context, program: bytes, entry_pc: int, initial_regs, cols, final_pc: int, final_regs, outputs,
                     compat_buffer: bool = False) -> bytes:
    """Everything `zigz prove` does after the VM has produced the trace (src/prover/prover.zig:91-226) + the "ZIGZ" v1
    serialization (src/prover/serialization.zig): returns the proof bytes."""
    c = _a64(cols)
    steps = c.shape[1] if c.ndim == 2 else 0
    ir, fr, out = _a64(initial_regs if initial_regs is not None else []), _a64(final_regs), _a64(outputs if outputs is not None else [])
    n = C.c_size_t(0)
    _keep, pp = _b8(program)
    args = (pp, len(program), entry_pc, _p64(ir) if ir.size else None, ir.size, _p64(c.reshape(-1)) if c.size else None, steps,
            final_pc, _p64(fr), _p64(out) if out.size else None, out.size, 1 if compat_buffer else 0)
    rc = lib().zh_prove_from_trace(ctx.handle, *args, None, 0, C.byref(n))
    if rc != -100:  # sizing call: OutOfMemory carries the exact size; anything else is the real error
        ctx.check(rc)
    buf = np.zeros(n.value, np.uint8)
    ctx.check(lib().zh_prove_from_trace(ctx.handle, *args, _p8(buf), buf.size, C.byref(n)))
    return buf[:n.value].tobytes()


def verify_proof(proof: bytes, program: bytes) -> str:
    """Verifier.verify (src/verifier/verifier.zig:49-294) on serialized proof bytes."""
    v = C.c_int32(-1)
    _k1, p1 = _b8(proof)
    _k2, p2 = _b8(program)
    rc = lib().zh_verify_proof(p1, len(proof), p2, len(program), C.byref(v))
    if rc != 0:
        raise ZigzError(rc)
    return VERDICTS[v.value]


# --------------------------------------------------------------------------- Lasso
TABLE_ADD, TABLE_XOR, TABLE_AND = 0, 1, 2


def _table_rows(bits: int, op: int) -> np.ndarray:
    """buildAddTable / buildXorTable / buildAndTable (src/lookups/table_builder.zig:126-213): rows (a, b, out),
    entry index = a * 2^bits + b.  Host-side table GENERATION is a caller convenience, not on the proving path."""
    m = 1 << bits
    a = np.repeat(np.arange(m, dtype=np.uint64), m)
    b = np.tile(np.arange(m, dtype=np.uint64), m)
    r = ((a + b) % np.uint64(m)) if op == TABLE_ADD else (a ^ b) if op == TABLE_XOR else (a & b)
    return np.stack([a % np.uint64(BABYBEAR_P), b % np.uint64(BABYBEAR_P), r % np.uint64(BABYBEAR_P)], axis=1)


def build_add_table(bits):
    return _table_rows(bits, TABLE_ADD)


def build_xor_table(bits):
    return _table_rows(bits, TABLE_XOR)


def build_and_table(bits):
    return _table_rows(bits, TABLE_AND)


@dataclass
class LassoProof:
    """src/lookups/lasso_prover.zig:27-62"""
    sumcheck_proof: SumcheckProof
    query_commitment: bytes
    table_commitment: bytes
    num_lookups: int


class LassoProver:
    """src/lookups/lasso_prover.zig:103-252. Tables / queries are (n, arity) uint64 rows = inputs || outputs."""

    @staticmethod
    def _run(ctx, call, n_queries):
        vmax = max(int(max(n_queries, 1) - 1).bit_length(), 1)
        rp, fp = np.zeros((vmax, 2), np.uint64), np.zeros(vmax, np.uint64)
        fe, nv = u64(0), u32(0)
        qc, tc = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
        ctx.check(call(_p64(rp), _p64(fp), C.byref(fe), C.byref(nv), _p8(qc), _p8(tc)))
        v = nv.value
        return LassoProof(SumcheckProof(v, rp[:v], fp[:v], fe.value), qc.tobytes(), tc.tobytes(), n_queries)

    @staticmethod
    def prove(ctx: Context, table, queries) -> LassoProof:  # :103-173
        t, q = _a64(table), _a64(queries)
        arity = t.shape[1] if t.ndim == 2 and t.size else (q.shape[1] if q.ndim == 2 else 1)
        nt = t.shape[0] if t.ndim == 2 else 0
        nq = q.shape[0] if q.ndim == 2 else 0
        tp = _p64(t.reshape(-1)) if t.size else None
        qp = _p64(q.reshape(-1)) if q.size else None
        return LassoProver._run(ctx, lambda *o: lib().zh_lasso_prove(ctx.handle, tp, nt, qp, nq, arity, *o), nq)

    @staticmethod
    def prove_with_mapping(ctx: Context, table, queries, mapping) -> LassoProof:  # :179-205
        t, q, m = _a64(table), _a64(queries), _a64(mapping)
        arity = t.shape[1]
        nq = q.shape[0] if q.ndim == 2 else 0
        qp = _p64(q.reshape(-1)) if q.size else None
        mp = _p64(m) if m.size else None
        return LassoProver._run(
            ctx, lambda *o: lib().zh_lasso_prove_with_mapping(ctx.handle, _p64(t.reshape(-1)), t.shape[0], qp, nq, mp, m.size,
                                                              arity, *o), nq)

    @staticmethod
    def prove_builtin(ctx: Context, op: int, bits: int, queries) -> LassoProof:
        q = _a64(queries)
        nq = q.shape[0]
        return LassoProver._run(ctx, lambda *o: lib().zh_lasso_prove_builtin(ctx.handle, op, bits, _p64(q.reshape(-1)), nq, *o), nq)

    @staticmethod
    def prove_builtin_batch(ctx: Context, jobs) -> List[LassoProof]:
        """jobs = [(op, bits, queries), ...]: the proofs `prove_builtin` returns one by one, with the sequential query
        commitments of all jobs running concurrently (zh_lasso_prove_builtin_batch). Raises the first job's error."""
        k = len(jobs)
        if k == 0:
            return []
        qs = [_a64(q) for _, _, q in jobs]
        nqs = [q.shape[0] if q.ndim == 2 else 0 for q in qs]
        vmax = [max(int(max(n, 1) - 1).bit_length(), 1) for n in nqs]
        rps = [np.zeros((v, 2), np.uint64) for v in vmax]
        fps = [np.zeros(v, np.uint64) for v in vmax]
        P64 = C.POINTER(u64)
        ops = (C.c_int32 * k)(*[int(op) for op, _, _ in jobs])
        bits = (u32 * k)(*[int(b) for _, b, _ in jobs])
        qptr = (P64 * k)(*[_p64(q.reshape(-1)) if q.size else P64() for q in qs])
        nq = (u64 * k)(*nqs)
        rpp = (P64 * k)(*[_p64(r) for r in rps])
        fpp = (P64 * k)(*[_p64(f) for f in fps])
        fe, nv, st = (u64 * k)(), (u32 * k)(), (C.c_int32 * k)()
        qc, tc = np.zeros((k, 32), np.uint8), np.zeros((k, 32), np.uint8)
        ctx.check(lib().zh_lasso_prove_builtin_batch(ctx.handle, k, ops, bits, qptr, nq, rpp, fpp, fe, nv, _p8(qc), _p8(tc), st))
        return [LassoProof(SumcheckProof(nv[j], rps[j][:nv[j]], fps[j][:nv[j]], fe[j]), qc[j].tobytes(), tc[j].tobytes(), nqs[j])
                for j in range(k)]
