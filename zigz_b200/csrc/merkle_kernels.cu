// sm_100a kernels for SimpleMerkleTree(BabyBear, SHA3Hasher) — /root/reference/src/commitments/merkle_tree.zig:283-400.
//   leaves : leaf_hashes[i] = SHA3-256(le64(values[i])), padded with SHA3-256(le64(0))      (:298-306)
//   levels : next[i] = SHA3-256(cur[2i] || cur[2i+1])                                        (:386-396)
//   open   : sibling gather from the retained levels (the reference recomputes them, :335-353)
// Level-synchronous, batched over the trees of one commit batch (blockIdx.y = tree). This path is bound by the
// integer ALU pipe (one Keccak-f[1600] per digest, ~4.3k LOP3/SHF each), not by HBM: each thread owns one
// permutation, state in 50 registers, no shared memory, no spills.
#include "keccak.cuh"
#include "kernels.h"

#include <cstdlib>

namespace zk {

#ifndef ZB_KECCAK_UNROLL
#define ZB_KECCAK_UNROLL 24
#endif
constexpr int KT = 128; // threads per CTA for the hashing kernels

__device__ __forceinline__ void store_digest(uint8_t *dst, const uint32_t (&d)[8]) {
    uint4 *o = reinterpret_cast<uint4 *>(dst);
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

// Share of the 29 rotations per round (24 rho + 5 theta) that run on the FMA pipe instead of the ALU pipe.
// MEASURED (profiles/r01_sweep1_tuning.txt, 2^24 leaves): 0 -> 8.38 ms, 8 lanes -> 9.01, 16 -> 10.09, 24 -> 11.19,
// all 29 -> 12.01 ms: IMAD.WIDE / IMAD.HI cost more issue slots than the two SHF they replace, so the all-ALU form
// stays the default; the variants are compiled only with -DZB_KECCAK_FMA_VARIANTS (ZB_KECCAK_V=1..4 selects).
#ifdef ZB_KECCAK_FMA_VARIANTS
constexpr uint32_t FMA_MASKS[5] = {0u, 0x1FEu, 0x1FFFEu, 0x1FFFFFEu, 0x3FFFFFFEu};
#endif
#ifndef ZB_KECCAK_DEFAULT_VARIANT
#define ZB_KECCAK_DEFAULT_VARIANT 0
#endif
// Rounds per loop iteration. Round 1 (profiles/r01_keccak_unroll.txt): leaves fully unrolled (24), nodes 8 per iteration
// (fully unrolled the code is 69 KB, more than the 32 KB L1.5 instruction cache). Round 2 (profiles/r02_sweep4.txt, 2^26
// leaves): the PEELED form 102 — round 0 and round 23 as straight-line code around a loop of 11 x 2 rounds (keccak.cuh) —
// gets the constant folding of the first round and the dead-lane elimination of the last one in ~12 KB of code:
// node levels 15.96 -> 15.37 ms, leaves 16.51 -> 15.12 ms, commit 32.71 -> 30.75 ms.
static int keccak_unroll(bool leaves) {
    static const int v_leaf = [] {
        const char *e = getenv("ZB_KECCAK_UNROLL_LEAF");
        return e && *e ? atoi(e) : 102;
    }();
    static const int v_node = [] {
        const char *e = getenv("ZB_KECCAK_UNROLL");
        return e && *e ? atoi(e) : 102;
    }();
    return leaves ? v_leaf : v_node;
}
static int keccak_variant() {
    static const int v = [] {
        const char *e = getenv("ZB_KECCAK_V");
        int x = e && *e ? atoi(e) : ZB_KECCAK_DEFAULT_VARIANT;
        return x < 0 || x > 4 ? 0 : x;
    }();
    return v;
}

void keccak_init_constants() {
    uint32_t pow2[33];
    for (int i = 0; i < 32; i++) pow2[i] = 1u << i;
    pow2[32] = 1u;
    cudaMemcpyToSymbol(keccak::POW2, pow2, sizeof(pow2));
}

// UR = Keccak rounds per loop iteration (24 = fully unrolled: ~69 KB of code, more than the 32 KB L1.5 instruction cache)
template <uint32_t FM, int UR = ZB_KECCAK_UNROLL>
__global__ void __launch_bounds__(KT) k_merkle_leaves(MerkleBatch b, uint64_t padded) {
    const uint32_t t = blockIdx.y;
    const uint32_t *vals = b.values[t];
    const uint64_t n = b.n_values[t];
    uint8_t *tree = b.tree[t];
    const uint64_t stride = (uint64_t)gridDim.x * KT;
    for (uint64_t i = (uint64_t)blockIdx.x * KT + threadIdx.x; i < padded; i += stride) {
        uint32_t v = i < n ? vals[i] : 0u;
        uint32_t d[8];
        keccak::sha3_leaf<UR, FM>(v, d);
        store_digest(tree + i * 32, d);
    }
}

template <uint32_t FM, int UR = ZB_KECCAK_UNROLL>
__device__ __forceinline__ void hash_pair(const uint8_t *in, uint8_t *out) {
    const uint4 *p = reinterpret_cast<const uint4 *>(in);
    uint4 a = p[0], b4 = p[1], c = p[2], e = p[3];
    uint32_t m[16] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w, c.x, c.y, c.z, c.w, e.x, e.y, e.z, e.w};
    uint32_t d[8];
    keccak::sha3_node<UR, FM>(m, d);
    store_digest(out, d);
}

template <uint32_t FM, int UR = ZB_KECCAK_UNROLL>
__global__ void __launch_bounds__(KT) k_merkle_level(MerkleBatch b, uint64_t in_off, uint64_t out_off, uint64_t width_out) {
    uint8_t *tree = b.tree[blockIdx.y];
    const uint8_t *in = tree + in_off * 32;
    uint8_t *out = tree + out_off * 32;
    const uint64_t stride = (uint64_t)gridDim.x * KT;
    for (uint64_t i = (uint64_t)blockIdx.x * KT + threadIdx.x; i < width_out; i += stride) hash_pair<FM, UR>(in + i * 64, out + i * 32);
}

// all remaining levels of one tree inside one CTA: width (<= MERKLE_TOP_WIDTH) digests at `level` down to the root
__global__ void __launch_bounds__(KT) k_merkle_top(MerkleBatch b, uint64_t padded, uint32_t level, uint64_t width) {
    uint8_t *tree = b.tree[blockIdx.x];
    while (width > 1) {
        const uint8_t *in = tree + (2 * padded - (2 * padded >> level)) * 32;
        uint8_t *out = tree + (2 * padded - (2 * padded >> (level + 1))) * 32;
        const uint64_t width_out = width / 2;
        for (uint64_t i = threadIdx.x; i < width_out; i += KT) hash_pair<0, 102>(in + i * 64, out + i * 32);
        __syncthreads(); // global writes of this CTA are visible to the CTA after the barrier
        width = width_out;
        level++;
    }
}

// single-CTA gathers into the host-mapped bulk area, then publish the mailbox sequence number
__device__ __forceinline__ void publish_seq(const Mailbox &mb) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        ((volatile unsigned long long *)mb.mail)[MAIL_WORDS] = mb.seq;
    }
}

__global__ void k_merkle_path(const uint8_t *tree, uint64_t padded, uint32_t height, uint64_t index, uint8_t *out, Mailbox mb) {
    // thread = (level, 16-byte half)
    const uint32_t l = threadIdx.x >> 1, half = threadIdx.x & 1;
    if (l < height) {
        const uint64_t sib = (index >> l) ^ 1ull;
        const uint64_t off = (2 * padded - (2 * padded >> l)) + sib;
        const uint4 *p = reinterpret_cast<const uint4 *>(tree + off * 32);
        reinterpret_cast<uint4 *>(out + (uint64_t)l * 32)[half] = p[half];
    }
    publish_seq(mb);
}

__global__ void k_merkle_roots(MerkleBatch b, uint64_t padded, uint8_t *out, Mailbox mb) {
    const uint32_t t = threadIdx.x >> 1, half = threadIdx.x & 1;
    if (t < b.count) {
        const uint4 *p = reinterpret_cast<const uint4 *>(b.tree[t] + (2 * padded - 2) * 32);
        reinterpret_cast<uint4 *>(out + (uint64_t)t * 32)[half] = p[half];
    }
    publish_seq(mb);
}

// `count` openings in one launch (Prover.generateCommitments opens 43 trees, prover.zig:420-443): CTA t gathers the path of
// leaf idx[t] of tree t into out + t * (height * 32), and the leaf value into vals[t]
__global__ void k_merkle_path_batch(const uint8_t *const *trees, const uint32_t *const *values, const uint64_t *idx, uint64_t padded,
                                    uint32_t height, uint8_t *out, uint32_t *vals) {
    const uint32_t t = blockIdx.x, l = threadIdx.x >> 1, half = threadIdx.x & 1;
    const uint64_t index = idx[t];
    if (l < height) {
        const uint64_t sib = (index >> l) ^ 1ull;
        const uint64_t off = (2 * padded - (2 * padded >> l)) + sib;
        const uint4 *p = reinterpret_cast<const uint4 *>(trees[t] + off * 32);
        reinterpret_cast<uint4 *>(out + ((uint64_t)t * height + l) * 32)[half] = p[half];
    }
    if (threadIdx.x == 0) vals[t] = values[t][index];
}
void launch_merkle_path_batch(const uint8_t *const *trees, const uint32_t *const *values, const uint64_t *idx, uint32_t count,
                              uint64_t padded, uint32_t height, uint8_t *out, uint32_t *vals, cudaStream_t st) {
    k_merkle_path_batch<<<count, 2 * 64, 0, st>>>(trees, values, idx, padded, height, out, vals);
}

static inline unsigned hash_grid(uint64_t items) {
    uint64_t g = (items + KT - 1) / KT;
    const uint64_t cap = 148ull * 64; // grid-stride beyond this
    if (g < 1) g = 1;
    return (unsigned)(g > cap ? cap : g);
}

void launch_merkle_leaves(const MerkleBatch &b, uint64_t padded, cudaStream_t st) {
    dim3 grid(hash_grid(padded), b.count);
    switch (keccak_variant()) {
#ifdef ZB_KECCAK_FMA_VARIANTS
    case 1: k_merkle_leaves<FMA_MASKS[1]><<<grid, KT, 0, st>>>(b, padded); break;
    case 2: k_merkle_leaves<FMA_MASKS[2]><<<grid, KT, 0, st>>>(b, padded); break;
    case 3: k_merkle_leaves<FMA_MASKS[3]><<<grid, KT, 0, st>>>(b, padded); break;
    case 4: k_merkle_leaves<FMA_MASKS[4]><<<grid, KT, 0, st>>>(b, padded); break;
#endif
    default:
        switch (keccak_unroll(true)) {
        case 102: k_merkle_leaves<0, 102><<<grid, KT, 0, st>>>(b, padded); break;
        case 111: k_merkle_leaves<0, 111><<<grid, KT, 0, st>>>(b, padded); break;
        case 12: k_merkle_leaves<0, 12><<<grid, KT, 0, st>>>(b, padded); break;
        case 8: k_merkle_leaves<0, 8><<<grid, KT, 0, st>>>(b, padded); break;
        case 6: k_merkle_leaves<0, 6><<<grid, KT, 0, st>>>(b, padded); break;
        case 4: k_merkle_leaves<0, 4><<<grid, KT, 0, st>>>(b, padded); break;
        case 2: k_merkle_leaves<0, 2><<<grid, KT, 0, st>>>(b, padded); break;
        default: k_merkle_leaves<0, 24><<<grid, KT, 0, st>>>(b, padded); break;
        }
        break;
    }
}

void launch_merkle_level(const MerkleBatch &b, uint64_t padded, uint32_t level, cudaStream_t st) {
    uint64_t width_out = padded >> (level + 1);
    dim3 grid(hash_grid(width_out), b.count);
    const uint64_t io = merkle_level_offset(padded, level), oo = merkle_level_offset(padded, level + 1);
    switch (keccak_variant()) {
#ifdef ZB_KECCAK_FMA_VARIANTS
    case 1: k_merkle_level<FMA_MASKS[1]><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
    case 2: k_merkle_level<FMA_MASKS[2]><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
    case 3: k_merkle_level<FMA_MASKS[3]><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
    case 4: k_merkle_level<FMA_MASKS[4]><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
#endif
    default:
        switch (keccak_unroll(false)) {
        case 102: k_merkle_level<0, 102><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        case 111: k_merkle_level<0, 111><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        case 12: k_merkle_level<0, 12><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        case 8: k_merkle_level<0, 8><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        case 6: k_merkle_level<0, 6><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        case 4: k_merkle_level<0, 4><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        case 2: k_merkle_level<0, 2><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        default: k_merkle_level<0, 24><<<grid, KT, 0, st>>>(b, io, oo, width_out); break;
        }
        break;
    }
}

void launch_merkle_top(const MerkleBatch &b, uint64_t padded, uint32_t level, cudaStream_t st) {
    k_merkle_top<<<b.count, KT, 0, st>>>(b, padded, level, padded >> level);
}

void launch_merkle_path(const uint8_t *tree, uint64_t padded, uint32_t height, uint64_t index, uint8_t *out, const Mailbox &mb,
                        cudaStream_t st) {
    k_merkle_path<<<1, 2 * 64, 0, st>>>(tree, padded, height, index, out, mb);
}

void launch_merkle_roots(const MerkleBatch &b, uint64_t padded, uint8_t *out, const Mailbox &mb, cudaStream_t st) {
    k_merkle_roots<<<1, 2 * MAX_BATCH, 0, st>>>(b, padded, out, mb);
}

} // namespace zk
