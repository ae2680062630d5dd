// sm_100a kernels for the BabyBear multilinear hot loops of zigz:
//   sumOverHypercube   /root/reference/src/poly/multilinear.zig:188-194
//   roundPolynomial    /root/reference/src/poly/multilinear.zig:205-232
//   partialEval        /root/reference/src/poly/multilinear.zig:154-180   (fused with the NEXT round's sums)
//   eval               /root/reference/src/poly/multilinear.zig:110-144   (O(N) LSB-first fold)
// All of them are HBM-bound streaming kernels: 128-bit coalesced loads, u64 per-thread accumulators
// (sums of < 2^33 canonical 31-bit values cannot overflow), warp-shuffle + shared-memory block reduction,
// one u64 atomicAdd per CTA and sum, and the last CTA to arrive publishes the canonical results to the
// host-mapped mailbox — no separate reduction kernel, no D2H copy.
#include "bb.cuh"
#include "kernels.h"

#include <atomic>
#include <cstdlib>
#include <type_traits>

namespace zk {

constexpr int THREADS = 256;
constexpr unsigned int TAIL_ABORT_TAG = 0xFFFFFFFFu; // challenge-word tag that makes polling kernels leave

// launch-shape knobs, overridable from the environment for tuning runs (read once)
static int tune(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Opt-in dynamic shared memory is a per-DEVICE function attribute: a process that drives several GPUs (zb_ctx_create_mask)
// must set it once on each of them.
#define ENSURE_DYN_SMEM(kernel, bytes)                                                                 \
    do {                                                                                               \
        static std::atomic<unsigned> done_{0};                                                         \
        int dev_ = 0;                                                                                  \
        cudaGetDevice(&dev_);                                                                          \
        if (!((done_.load(std::memory_order_acquire) >> (dev_ & 31)) & 1u)) {                          \
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);          \
            done_.fetch_or(1u << (dev_ & 31), std::memory_order_release);                              \
        }                                                                                              \
    } while (0)

static inline int grid_for(uint64_t work_items, int sm_count, int ctas_per_sm) {
    uint64_t need = (work_items + THREADS - 1) / THREADS;
    uint64_t cap = (uint64_t)sm_count * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Cross-GPU sum of NS canonical words, executed by warp 0 of the last CTA of the kernel on EVERY rank at the same
// point of the same round: lane q stores this rank's words into rank q's exchange buffer (NVLink P2P store) and
// raises this rank's flag there; then lane q waits for rank q's flag in the LOCAL buffer and reads its words; a warp
// shuffle adds the rows. No NCCL call, no extra launch: the reduction rides in the tail of the kernel that produced
// the partial sums. Returns false when a peer's row has not arrived after xv->patience cycles (default ~30 s: ranks
// that upload their own shards first can be seconds apart; ZB_XCHG_PATIENCE_S).
template <int NS>
__device__ __forceinline__ bool xchg_allreduce(unsigned long long (&tot)[NS], const XchgView *xv, unsigned long long seq) {
    const int lane = threadIdx.x & 31;
    const int world = xv->world, rank = xv->rank;
    const unsigned long long set = (seq & 1ull) * XCHG_SET_WORDS;
    // Self-validating words (the payload values are canonical field elements < 2^31): (low 32 bits of seq) << 32 | value. No
    // system-scope fence between payload and flag, no flag at all: a row has arrived when every one of its words carries the
    // tag. An 8-byte peer store is atomic, and the alternating sets keep round seq - 1 apart from round seq.
    if (lane < world) {
        volatile unsigned long long *dst = xv->peer[lane] + set;
#pragma unroll
        for (int k = 0; k < NS; k++) dst[rank * XCHG_ROW + k] = mail_tagged(seq, tot[k]);
    }
    bool ok = true;
    unsigned long long v[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) v[k] = 0;
    if (lane < world) {
        volatile unsigned long long *mine = xv->peer[rank] + set;
        const long long t0 = clock64(), patience = xv->patience;
        const unsigned long long tag = seq & 0xffffffffull;
        for (;;) {
            bool all = true;
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const unsigned long long w = mine[lane * XCHG_ROW + k];
                all = all && (w >> 32) == tag;
                v[k] = w & 0xffffffffull;
            }
            if (all) break;
            if (clock64() - t0 > patience) {
                ok = false;
                break;
            }
        }
        if (xv->stats != nullptr) { // the slowest peer's row sets this rank's wait
            long long waited = clock64() - t0;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const long long other = __shfl_xor_sync(__activemask(), waited, o, 16);
                waited = other > waited ? other : waited;
            }
            if (lane == 0) {
                atomicAdd(&xv->stats[0], (unsigned long long)waited);
                atomicAdd(&xv->stats[1], 1ull);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NS; k++) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        tot[k] = bb::reduce64_scaled<0>(x); // sums of <= 16 canonical words: exact
    }
    return __all_sync(0xffffffffu, ok);
}

// Block-reduce NS u64 partial sums, add them to the global accumulators; the last CTA converts the totals with
// `fin` (canonical field elements), optionally sums them over all GPUs (mb.xchg) and publishes payload + sequence
// number to the mailbox.
template <int NS, typename Fin, int TPB = THREADS>
__device__ __forceinline__ void publish_sums(unsigned long long (&s)[NS], const Mailbox &mb, Fin fin) {
    __shared__ unsigned long long sm[NS][TPB / 32];
    __shared__ unsigned long long s_tot[NS];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; k++) {
        unsigned long long v = warp_sum(s[k]);
        if (lane == 0) sm[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) {
            unsigned long long v = 0;
#pragma unroll
            for (int w = 0; w < TPB / 32; w++) v += sm[k][w];
            tot[k] = v;
        }
        if (gridDim.x == 1) {
            // a single CTA already holds the totals: no accumulators, no ticket, no device-scope fences (small tables are
            // pure latency: ~3 us of dependent global atomics per launch)
            is_last = true;
        } else {
#pragma unroll
            for (int k = 0; k < NS; k++) atomicAdd(&mb.acc[k], tot[k]);
            __threadfence();
            unsigned int t = atomicAdd(mb.ticket, 1u);
            is_last = (t == gridDim.x - 1);
            if (is_last) {
                __threadfence();
#pragma unroll
                for (int k = 0; k < NS; k++) tot[k] = atomicExch(&mb.acc[k], 0ull); // read + re-arm
                *mb.ticket = 0u;
            }
        }
        if (is_last) {
            fin(tot);
            if (mb.xchg == nullptr) {
                if (mb.tagged) { // self-validating words: no fence
#pragma unroll
                    for (int k = 0; k < NS; k++) ((volatile unsigned long long *)mb.mail)[k] = mail_tagged(mb.seq, tot[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < NS; k++) ((volatile unsigned long long *)mb.mail)[k] = tot[k];
                    __threadfence_system();
                }
                ((volatile unsigned long long *)mb.mail)[MAIL_WORDS] = mb.seq;
            } else {
#pragma unroll
                for (int k = 0; k < NS; k++) s_tot[k] = tot[k];
            }
        }
    }
    if (mb.xchg != nullptr) { // uniform across the grid: every CTA takes the barrier, only the last one exchanges
        __syncthreads();
        if (is_last && warp == 0) {
            unsigned long long tot[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) tot[k] = s_tot[k];
            const bool ok = xchg_allreduce<NS>(tot, mb.xchg, mb.xseq);
            if (lane == 0) {
                volatile unsigned long long *mail = (volatile unsigned long long *)mb.mail;
                if (!ok) { // status word: 1 = a peer never arrived (the host clears it); ordered before the payload it qualifies
                    mail[MAIL_WORDS - 1] = 1ull;
                    __threadfence_system();
                }
                // tagged payload (mb.tagged is set for these mailboxes): the host waits for the tags, no fence on the way
#pragma unroll
                for (int k = 0; k < NS; k++) mail[k] = mail_tagged(mb.seq, tot[k]);
                mail[MAIL_WORDS] = mb.seq;
            }
        }
    }
}

template <int D>
struct NSums {
    static constexpr int value = D == 1 ? 2 : D + 1;
};

// accumulate the round-polynomial evaluations contributed by one MSB-first pair (lo_k, hi_k), k < D
template <int D>
__device__ __forceinline__ void accum_pair(const uint32_t (&lo)[D], const uint32_t (&hi)[D],
                                           unsigned long long (&s)[NSums<D>::value]) {
    if constexpr (D == 1) {
        s[0] += lo[0];
        s[1] += hi[0];
    } else if constexpr (D == 2) {
        // lazy Montgomery products (< 2P) go straight into the u64 accumulators: < 2^31 terms of < 2^32 cannot overflow
        uint32_t d0 = bb::sub_lazy(hi[0], lo[0]), d1 = bb::sub(hi[1], lo[1]);
        s[0] += bb::mont_mul_lazy(lo[0], lo[1]);
        s[1] += bb::mont_mul_lazy(hi[0], hi[1]);
        s[2] += bb::mont_mul_lazy(d0, d1);
    } else {
        uint32_t d0 = bb::sub(hi[0], lo[0]), d1 = bb::sub(hi[1], lo[1]), d2 = bb::sub(hi[2], lo[2]);
        // value at X = -1: lo - d; the first factor may stay lazy, the other two must be canonical (mont_mul_lazy)
        uint32_t m0 = bb::sub_lazy(lo[0], d0), m1 = bb::sub(lo[1], d1), m2 = bb::sub(lo[2], d2);
        s[0] += bb::mont_mul_lazy(bb::mont_mul_lazy(lo[0], lo[1]), lo[2]);
        s[1] += bb::mont_mul_lazy(bb::mont_mul_lazy(hi[0], hi[1]), hi[2]);
        s[2] += bb::mont_mul_lazy(bb::mont_mul_lazy(m0, m1), m2);
        s[3] += bb::mont_mul_lazy(bb::mont_mul_lazy(d0, d1), d2);
    }
}

// raw u64 totals -> canonical field elements (undoing the Montgomery factors R^-(D-1))
template <int D>
struct Finish {
    __device__ void operator()(unsigned long long (&t)[NSums<D>::value]) const {
#pragma unroll
        for (int k = 0; k < NSums<D>::value; k++) t[k] = bb::reduce64_scaled<D - 1>(t[k]);
    }
};

#define UNPACK4(v, a) \
    {                 \
        a[0] = v.x;   \
        a[1] = v.y;   \
        a[2] = v.z;   \
        a[3] = v.w;   \
    }

// ---------------------------------------------------------------------------------------------
// round sums only (first round): pairs (i, i + h); h4 = h / 4 uint4 per half
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(THREADS) k_round_sums_v4(PolySet ps, uint64_t h4, Mailbox mb) {
    constexpr int NS = NSums<D>::value;
    unsigned long long s[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) s[k] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * THREADS;
#pragma unroll(D == 1 ? 4 : 2)
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < h4; i += stride) {
        uint32_t lo[D][4], hi[D][4];
#pragma unroll
        for (int k = 0; k < D; k++) {
            const uint4 *p = reinterpret_cast<const uint4 *>(ps.src[k]);
            uint4 a = p[i], b = p[i + h4];
            UNPACK4(a, lo[k]);
            UNPACK4(b, hi[k]);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t l[D], h[D];
#pragma unroll
            for (int k = 0; k < D; k++) {
                l[k] = lo[k][c];
                h[k] = hi[k][c];
            }
            accum_pair<D>(l, h, s);
        }
    }
    publish_sums<NS>(s, mb, Finish<D>());
}

// scalar variant for tiny / unaligned sizes: h pairs
template <int D>
__global__ void __launch_bounds__(THREADS) k_round_sums_s(PolySet ps, uint64_t h, Mailbox mb) {
    constexpr int NS = NSums<D>::value;
    unsigned long long s[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) s[k] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < h; i += stride) {
        uint32_t l[D], hh[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            l[k] = ps.src[k][i];
            hh[k] = ps.src[k][i + h];
        }
        accum_pair<D>(l, hh, s);
    }
    publish_sums<NS>(s, mb, Finish<D>());
}

// ---------------------------------------------------------------------------------------------
// fold + next-round sums. Current length n = 4q. Thread handles i in [0, q):
//   new[i]     = lerp(e[i],     e[i + 2q])      (pair (i, i + n/2))
//   new[i + q] = lerp(e[i + q], e[i + 3q])
// and (new[i], new[i + q]) is exactly the next round's MSB-first pair. q4 = q / 4.
// In place is safe: a thread only touches indices congruent to its own i modulo q.
// ---------------------------------------------------------------------------------------------
// pre-launched kernels: obtain the challenge (see ChalSrc). Returns false when the kernel must leave untouched.
__device__ __forceinline__ bool acquire_challenge(const ChalSrc &cs, uint32_t &r, uint32_t &rp) {
    __shared__ uint32_t s_r, s_rp;
    __shared__ int s_ab;
    if (threadIdx.x == 0) {
        unsigned long long w = 0;
        int ab = 0;
        if (atomicExch(cs.claim, cs.tag) != cs.tag) { // first CTA of this launch: talk to the host
            const long long t0 = clock64();
            for (;;) {
                w = *(const volatile unsigned long long *)cs.chal;
                const unsigned int tag = (unsigned int)(w >> 32);
                if (tag == cs.tag) break;
                if (tag == TAIL_ABORT_TAG || clock64() - t0 > 400000000ll) {
                    ab = 1;
                    break;
                }
            }
            if (ab) w = ((unsigned long long)cs.tag << 32) | 0x80000000ull; // bit 31 (never set in a challenge) = leave
            *(volatile unsigned long long *)cs.bcast = w;
            __threadfence();
        } else {
            for (;;) {
                w = *(const volatile unsigned long long *)cs.bcast;
                if ((unsigned int)(w >> 32) == cs.tag) break;
            }
            ab = (w & 0x80000000ull) != 0;
        }
        s_ab = ab;
        s_r = (uint32_t)w & 0x7FFFFFFFu;
        s_rp = bb::shoup_pre(s_r);
    }
    __syncthreads();
    r = s_r;
    rp = s_rp;
    return !s_ab;
}

template <int D, int U>
__global__ void __launch_bounds__(THREADS) k_fold_sums_v4(PolySet ps, uint64_t q4, uint32_t r, uint32_t rp, Mailbox mb, ChalSrc cs) {
    if (cs.chal != nullptr && !acquire_challenge(cs, r, rp)) return;
    constexpr int NS = NSums<D>::value;
    unsigned long long s[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) s[k] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * THREADS;
    for (uint64_t i0 = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i0 < q4; i0 += U * stride) {
        // issue every load of the U independent items first: src may alias dst (in place), so the compiler
        // cannot hoist the next item's loads above this item's stores on its own
        uint4 v[U][D][4];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = i0 + u * stride;
            if (i < q4) {
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const uint4 *p = reinterpret_cast<const uint4 *>(ps.src[k]);
                    v[u][k][0] = p[i];
                    v[u][k][1] = p[i + q4];
                    v[u][k][2] = p[i + 2 * q4];
                    v[u][k][3] = p[i + 3 * q4];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = i0 + u * stride;
            if (i < q4) {
                uint32_t nlo[D][4], nhi[D][4];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    uint32_t e0[4], e1[4], e2[4], e3[4];
                    UNPACK4(v[u][k][0], e0);
                    UNPACK4(v[u][k][1], e1);
                    UNPACK4(v[u][k][2], e2);
                    UNPACK4(v[u][k][3], e3);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        nlo[k][c] = bb::lerp(e0[c], e2[c], r, rp);
                        nhi[k][c] = bb::lerp(e1[c], e3[c], r, rp);
                    }
                    uint4 *o = reinterpret_cast<uint4 *>(ps.dst[k]);
                    o[i] = make_uint4(nlo[k][0], nlo[k][1], nlo[k][2], nlo[k][3]);
                    o[i + q4] = make_uint4(nhi[k][0], nhi[k][1], nhi[k][2], nhi[k][3]);
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t l[D], h[D];
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        l[k] = nlo[k][c];
                        h[k] = nhi[k][c];
                    }
                    accum_pair<D>(l, h, s);
                }
            }
        }
    }
    publish_sums<NS>(s, mb, Finish<D>());
}

// scalar variant, q >= 1 quarter length
template <int D>
__global__ void __launch_bounds__(THREADS) k_fold_sums_s(PolySet ps, uint64_t q, uint32_t r, uint32_t rp, Mailbox mb) {
    constexpr int NS = NSums<D>::value;
    unsigned long long s[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) s[k] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < q; i += stride) {
        uint32_t l[D], h[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            const uint32_t *p = ps.src[k];
            uint32_t e0 = p[i], e1 = p[i + q], e2 = p[i + 2 * q], e3 = p[i + 3 * q];
            l[k] = bb::lerp(e0, e2, r, rp);
            h[k] = bb::lerp(e1, e3, r, rp);
            ps.dst[k][i] = l[k];
            ps.dst[k][i + q] = h[k];
        }
        accum_pair<D>(l, h, s);
    }
    publish_sums<NS>(s, mb, Finish<D>());
}

// last fold (n == 2): one value per polynomial remains; payload = the D final evaluations
template <int D>
__global__ void k_fold_last(PolySet ps, uint32_t r, uint32_t rp, Mailbox mb) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < D; k++) {
            uint32_t v = bb::lerp(ps.src[k][0], ps.src[k][1], r, rp);
            ps.dst[k][0] = v;
            ((volatile unsigned long long *)mb.mail)[k] = mb.tagged ? mail_tagged(mb.seq, v) : v;
        }
        if (!mb.tagged) __threadfence_system();
        ((volatile unsigned long long *)mb.mail)[MAIL_WORDS] = mb.seq;
    }
}

// ---------------------------------------------------------------------------------------------
// fold (0..2 variables) + bivariate round grid of the NEXT two variables (see launch_fold_grid in kernels.h)
// ---------------------------------------------------------------------------------------------
template <int D>
struct NPts {
    static constexpr int value = D == 1 ? 2 : D + 1;
};

// values of the line through (0 -> u, 1 -> v) on the point set of degree D: {0,1}, {0,1,inf}, {0,1,-1,inf}
template <int D>
__device__ __forceinline__ void expand_line(uint32_t u, uint32_t v, uint32_t (&o)[NPts<D>::value]) {
    o[0] = u;
    o[1] = v;
    if constexpr (D == 2) o[2] = bb::sub(v, u);
    if constexpr (D == 3) {
        const uint32_t dd = bb::sub(v, u);
        o[2] = bb::sub(u, dd);
        o[3] = dd;
    }
}

template <int D, int NS>
struct FinishGrid {
    __device__ void operator()(unsigned long long (&t)[NS]) const {
#pragma unroll
        for (int k = 0; k < NS; k++) t[k] = bb::reduce64_scaled<D - 1>(t[k]);
    }
};

// Kernel arguments of the fold step: one folded variable takes (r1, shoup(r1), r2, shoup(r2)) and interpolates; two folded
// variables take the four bilinear weights in Montgomery form and bind both in one dot product (bb::dot4).
struct FoldArgs {
    uint32_t a[4];
};
template <int FV>
static FoldArgs fold_args(uint32_t r1, uint32_t r2) {
    FoldArgs f;
    if (FV == 2) {
        const bb::BilinearWeights w = bb::bilinear_weights(r1, r2);
        // quarter t = 2 b1 + b2 (b1 = the variable bound by r1): weights in the order a[j][0..3]
        f.a[0] = w.w[0], f.a[1] = w.w[1], f.a[2] = w.w[2], f.a[3] = w.w[3];
    } else {
        f.a[0] = r1, f.a[1] = bb::shoup_pre(r1), f.a[2] = r2, f.a[3] = bb::shoup_pre(r2);
    }
    return f;
}

// Plain-load version (used below 2^16 entries, where the pass is not bandwidth-bound): 8-byte loads, one polynomial at a
// time. (A 4-byte-load variant with all polynomials prefetched was measured at 2.3-3.2 TB/s and dropped, r01_grid_sweep.txt.)
template <int D, int FV, int VEC>
__global__ void __launch_bounds__(THREADS) k_fold_grid(PolySet ps, uint64_t mq, uint32_t r1, uint32_t rp1, uint32_t r2, uint32_t rp2,
                                                       Mailbox mb) {
    constexpr int NP = NPts<D>::value, NS = NP * NP, NT = 1 << FV;
    unsigned long long s[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) s[k] = 0;
    // vector units: the folded tables have m elements = 4 mq vectors of VEC elements; an input table is NT m long
    const uint64_t stride = (uint64_t)gridDim.x * THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < mq; i += stride) {
        uint32_t f[D][4][VEC]; // [poly][quarter (b1 b2)][vector lane]
        static_assert(VEC == 2, "8-byte vectors");
        {
#pragma unroll
            for (int k = 0; k < D; k++) {
                const uint2 *p = reinterpret_cast<const uint2 *>(ps.src[k]);
                uint2 a[4][NT];
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int t = 0; t < NT; t++) a[j][t] = p[i + j * mq + t * (4 * mq)];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if constexpr (FV == 0) {
                        f[k][j][0] = a[j][0].x;
                        f[k][j][1] = a[j][0].y;
                    } else if constexpr (FV == 1) {
                        f[k][j][0] = bb::lerp(a[j][0].x, a[j][1].x, r1, rp1);
                        f[k][j][1] = bb::lerp(a[j][0].y, a[j][1].y, r1, rp1);
                    } else {
                        // FV == 2: (r1, rp1, r2, rp2) carry the four bilinear weights (fold_weights)
                        f[k][j][0] = bb::dot4(a[j][0].x, a[j][1].x, a[j][2].x, a[j][3].x, r1, rp1, r2, rp2);
                        f[k][j][1] = bb::dot4(a[j][0].y, a[j][1].y, a[j][2].y, a[j][3].y, r1, rp1, r2, rp2);
                    }
                }
                if constexpr (FV > 0) {
                    uint2 *o = reinterpret_cast<uint2 *>(ps.dst[k]);
#pragma unroll
                    for (int j = 0; j < 4; j++) o[i + j * mq] = make_uint2(f[k][j][0], f[k][j][1]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < VEC; c++) {
            uint32_t val[D][NP][NP];
#pragma unroll
            for (int k = 0; k < D; k++) {
                uint32_t x0[NP], x1[NP]; // the X line at Y = 0 (quarters 0 -> 2) and at Y = 1 (quarters 1 -> 3)
                expand_line<D>(f[k][0][c], f[k][2][c], x0);
                expand_line<D>(f[k][1][c], f[k][3][c], x1);
#pragma unroll
                for (int ix = 0; ix < NP; ix++) expand_line<D>(x0[ix], x1[ix], val[k][ix]);
            }
#pragma unroll
            for (int ix = 0; ix < NP; ix++)
#pragma unroll
                for (int iy = 0; iy < NP; iy++) {
                    if constexpr (D == 1) s[ix * NP + iy] += val[0][ix][iy];
                    else if constexpr (D == 2) s[ix * NP + iy] += bb::mont_mul_lazy(val[0][ix][iy], val[1][ix][iy]);
                    else s[ix * NP + iy] += bb::mont_mul_lazy(bb::mont_mul_lazy(val[0][ix][iy], val[1][ix][iy]), val[2][ix][iy]);
                }
        }
    }
    publish_sums<NS>(s, mb, FinishGrid<D, NS>());
}

// The same pass with the loads decoupled from the register file: every thread streams ITS OWN operands through a
// private ring of shared-memory slots with cp.async (LDGSTS, 8 bytes per slot), STAGES iterations deep. No barrier is
// needed (a thread only ever reads the slots it filled), the bytes in flight per SM are set by the ring (~100-150 KB)
// instead of by registers x occupancy, which is what the register-heavy grid arithmetic could not provide.
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int D, int FV, int STAGES_, int TPB>
struct GridAsync {
    static constexpr int NT = 1 << FV;
    static constexpr int NSLOT = D * 4 * NT;
    static constexpr int STAGE_BYTES = NSLOT * TPB * 8;
    static constexpr int STAGES = STAGES_;
    static constexpr int SMEM = STAGES * STAGE_BYTES;
    static_assert(STAGES >= 2 && SMEM <= 222 * 1024, "ring depth");
};

// DOT (FV == 1 only): (r1, rp1) carry the Montgomery weights ((1-r) R, r R) and the variable is bound as a two-term dot
// product (bb::dot2) instead of an interpolation with a Shoup product (ZB_GRID_DOT2)
template <int D, int FV, int STAGES_, int TPB, bool DOT = false>
__global__ void __launch_bounds__(TPB) k_fold_grid_async(PolySet ps, uint64_t mq, uint32_t r1, uint32_t rp1, uint32_t r2,
                                                             uint32_t rp2, Mailbox mb) {
    using G = GridAsync<D, FV, STAGES_, TPB>;
    constexpr int NP = NPts<D>::value, NS = NP * NP, NT = G::NT, NSLOT = G::NSLOT, STAGES = G::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2 *ring = reinterpret_cast<uint2 *>(smem_raw); // [STAGES][NSLOT][THREADS]
    unsigned long long s[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) s[k] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * TPB;
    const uint64_t i0 = (uint64_t)blockIdx.x * TPB + threadIdx.x;
    const uint64_t cnt = i0 < mq ? (mq - i0 + stride - 1) / stride : 0;
    auto issue = [&](int stage, uint64_t i) {
#pragma unroll
        for (int k = 0; k < D; k++) {
            const uint2 *p = reinterpret_cast<const uint2 *>(ps.src[k]);
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int t = 0; t < NT; t++)
                    cp_async8(&ring[((size_t)stage * NSLOT + (k * 4 + j) * NT + t) * TPB + threadIdx.x], p + i + j * mq + t * (4 * mq));
        }
    };
#pragma unroll
    for (int st = 0; st < STAGES - 1; st++) {
        if ((uint64_t)st < cnt) issue(st, i0 + st * stride);
        cp_async_commit();
    }
    for (uint64_t it = 0; it < cnt; it++) {
        const uint64_t nxt = it + STAGES - 1;
        if (nxt < cnt) issue((int)(nxt % STAGES), i0 + nxt * stride);
        cp_async_commit();
        cp_async_wait<STAGES - 1>(); // everything but the newest STAGES-1 groups has landed: iteration `it` is in its slots
        const int stage = (int)(it % STAGES);
        const uint64_t i = i0 + it * stride;
        uint32_t f[D][4][2];
#pragma unroll
        for (int k = 0; k < D; k++) {
            uint2 a[4][NT];
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int t = 0; t < NT; t++) a[j][t] = ring[((size_t)stage * NSLOT + (k * 4 + j) * NT + t) * TPB + threadIdx.x];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if constexpr (FV == 0) {
                    f[k][j][0] = a[j][0].x;
                    f[k][j][1] = a[j][0].y;
                } else if constexpr (FV == 1) {
                    if constexpr (DOT) {
                        f[k][j][0] = bb::dot2(a[j][0].x, a[j][1].x, r1, rp1);
                        f[k][j][1] = bb::dot2(a[j][0].y, a[j][1].y, r1, rp1);
                    } else {
                        f[k][j][0] = bb::lerp(a[j][0].x, a[j][1].x, r1, rp1);
                        f[k][j][1] = bb::lerp(a[j][0].y, a[j][1].y, r1, rp1);
                    }
                } else {
                    // FV == 2: (r1, rp1, r2, rp2) carry the four bilinear weights (fold_weights)
                    f[k][j][0] = bb::dot4(a[j][0].x, a[j][1].x, a[j][2].x, a[j][3].x, r1, rp1, r2, rp2);
                    f[k][j][1] = bb::dot4(a[j][0].y, a[j][1].y, a[j][2].y, a[j][3].y, r1, rp1, r2, rp2);
                }
            }
            if constexpr (FV > 0) {
                uint2 *o = reinterpret_cast<uint2 *>(ps.dst[k]);
#pragma unroll
                for (int j = 0; j < 4; j++) o[i + j * mq] = make_uint2(f[k][j][0], f[k][j][1]);
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            uint32_t val[D][NP][NP];
#pragma unroll
            for (int k = 0; k < D; k++) {
                uint32_t x0[NP], x1[NP];
                expand_line<D>(f[k][0][c], f[k][2][c], x0);
                expand_line<D>(f[k][1][c], f[k][3][c], x1);
#pragma unroll
                for (int ix = 0; ix < NP; ix++) expand_line<D>(x0[ix], x1[ix], val[k][ix]);
            }
#pragma unroll
            for (int ix = 0; ix < NP; ix++)
#pragma unroll
                for (int iy = 0; iy < NP; iy++) {
                    if constexpr (D == 1) s[ix * NP + iy] += val[0][ix][iy];
                    else if constexpr (D == 2) s[ix * NP + iy] += bb::mont_mul_lazy(val[0][ix][iy], val[1][ix][iy]);
                    else s[ix * NP + iy] += bb::mont_mul_lazy(bb::mont_mul_lazy(val[0][ix][iy], val[1][ix][iy]), val[2][ix][iy]);
                }
        }
    }
    cp_async_wait<0>();
    publish_sums<NS, FinishGrid<D, NS>, TPB>(s, mb, FinishGrid<D, NS>());
}

// ---------------------------------------------------------------------------------------------
// Bulk asynchronous copies (cp.async.bulk, SASS UBLKCP) completing on shared-memory mbarriers (SYNCS): ONE instruction
// moves a KB-sized contiguous slab global -> shared without touching the register file or the LSU issue slots of the
// consumer warps. The streaming kernels below use them as a producer/consumer ring: a `full` barrier per stage carries
// the expected byte count (expect_tx), an `empty` barrier hands the stage back to the producer.
// ---------------------------------------------------------------------------------------------
namespace blk {
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { // arrives once and arms the byte count
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(saddr(bar)),
        "r"(parity)
        : "memory");
}
// bytes: multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(saddr(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
} // namespace blk

// k_fold_grid_async with the per-thread LDGSTS ring (24-48 8-byte copies per thread and iteration) replaced by a CTA-wide
// ring of bulk copies: a tile is TPB consecutive vector units; each of the D*4*2^FV operand streams of the tile is one
// contiguous TPB*8-byte slab, fetched by ONE cp.async.bulk issued from the producer warp. The consumer warps wait on the
// stage's `full` mbarrier, read their operands with conflict-free 8-byte LDS and hand the stage back through `empty`.
template <int D, int FV, int STAGES_, int TPB>
struct GridBulk {
    static constexpr int NT = 1 << FV;
    static constexpr int NSLOT = D * 4 * NT;
    static constexpr int SLAB_BYTES = TPB * 8;
    static constexpr int STAGE_BYTES = NSLOT * SLAB_BYTES;
    static constexpr int STAGES = STAGES_;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128; // + barriers (+ alignment slack)
    static constexpr int THREADS_ALL = TPB + 32; // consumers + one producer warp
    static_assert(STAGES >= 2 && SMEM <= 227 * 1024, "ring depth");
};

template <int D, int FV, int STAGES_, int TPB>
__global__ void __launch_bounds__(TPB + 32, 1) k_fold_grid_bulk(PolySet ps, uint64_t mq, uint32_t r1, uint32_t rp1, uint32_t r2,
                                                                uint32_t rp2, Mailbox mb) {
    using G = GridBulk<D, FV, STAGES_, TPB>;
    constexpr int NP = NPts<D>::value, NS = NP * NP, NT = G::NT, NSLOT = G::NSLOT, STAGES = G::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2 *ring = reinterpret_cast<uint2 *>(smem_raw); // [STAGES][NSLOT][TPB]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * G::STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int st = 0; st < STAGES; st++) {
            blk::mbar_init(&full[st], 1);          // the producer's arrive.expect_tx
            blk::mbar_init(&empty[st], TPB / 32);  // one arrival per consumer warp
        }
        blk::fence_barrier_init();
    }
    __syncthreads();
    const uint64_t n_tiles = mq / TPB; // mq is a multiple of TPB (launcher)
    const uint64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    unsigned long long s[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) s[k] = 0;
    if (tid >= TPB) {
        // ---- producer warp ----
        const int lane = tid - TPB;
        for (uint64_t it = 0; it < my_tiles; it++) {
            const int stage = (int)(it % STAGES);
            const uint32_t use = (uint32_t)(it / STAGES);
            if (use > 0) blk::mbar_wait(&empty[stage], (use - 1) & 1); // the consumers have drained the previous tile of this stage
            if (lane == 0) blk::mbar_expect_tx(&full[stage], G::STAGE_BYTES);
            __syncwarp();
            const uint64_t i0 = (blockIdx.x + it * gridDim.x) * TPB;
            for (int slot = lane; slot < NSLOT; slot += 32) {
                const int k = slot / (4 * NT), j = (slot / NT) % 4, t = slot % NT;
                const uint2 *src = reinterpret_cast<const uint2 *>(ps.src[k]) + i0 + (uint64_t)j * mq + (uint64_t)t * (4 * mq);
                blk::g2s(&ring[((size_t)stage * NSLOT + slot) * TPB], src, G::SLAB_BYTES, &full[stage]);
            }
        }
    } else {
        // ---- consumer warps ----
        for (uint64_t it = 0; it < my_tiles; it++) {
            const int stage = (int)(it % STAGES);
            blk::mbar_wait(&full[stage], (uint32_t)(it / STAGES) & 1);
            const uint64_t i = (blockIdx.x + it * gridDim.x) * TPB + tid;
            uint32_t f[D][4][2];
#pragma unroll
            for (int k = 0; k < D; k++) {
                uint2 a[4][NT];
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int t = 0; t < NT; t++) a[j][t] = ring[((size_t)stage * NSLOT + (k * 4 + j) * NT + t) * TPB + tid];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if constexpr (FV == 0) {
                        f[k][j][0] = a[j][0].x;
                        f[k][j][1] = a[j][0].y;
                    } else if constexpr (FV == 1) {
                        f[k][j][0] = bb::lerp(a[j][0].x, a[j][1].x, r1, rp1);
                        f[k][j][1] = bb::lerp(a[j][0].y, a[j][1].y, r1, rp1);
                    } else {
                        f[k][j][0] = bb::dot4(a[j][0].x, a[j][1].x, a[j][2].x, a[j][3].x, r1, rp1, r2, rp2);
                        f[k][j][1] = bb::dot4(a[j][0].y, a[j][1].y, a[j][2].y, a[j][3].y, r1, rp1, r2, rp2);
                    }
                }
                if (k == D - 1) { // every operand of this tile is in registers: hand the stage back before the arithmetic
                    __syncwarp();
                    if ((tid & 31) == 0) blk::mbar_arrive(&empty[stage]);
                }
                if constexpr (FV > 0) {
                    uint2 *o = reinterpret_cast<uint2 *>(ps.dst[k]);
#pragma unroll
                    for (int j = 0; j < 4; j++) o[i + j * mq] = make_uint2(f[k][j][0], f[k][j][1]);
                }
            }
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t val[D][NP][NP];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    uint32_t x0[NP], x1[NP];
                    expand_line<D>(f[k][0][c], f[k][2][c], x0);
                    expand_line<D>(f[k][1][c], f[k][3][c], x1);
#pragma unroll
                    for (int ix = 0; ix < NP; ix++) expand_line<D>(x0[ix], x1[ix], val[k][ix]);
                }
#pragma unroll
                for (int ix = 0; ix < NP; ix++)
#pragma unroll
                    for (int iy = 0; iy < NP; iy++) {
                        if constexpr (D == 1) s[ix * NP + iy] += val[0][ix][iy];
                        else if constexpr (D == 2) s[ix * NP + iy] += bb::mont_mul_lazy(val[0][ix][iy], val[1][ix][iy]);
                        else s[ix * NP + iy] += bb::mont_mul_lazy(bb::mont_mul_lazy(val[0][ix][iy], val[1][ix][iy]), val[2][ix][iy]);
                    }
            }
        }
    }
    publish_sums<NS, FinishGrid<D, NS>, TPB + 32>(s, mb, FinishGrid<D, NS>());
}

template <int D, int FV, int STAGES_, int TPB>
static void fold_grid_bulk_launch_s(const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm, cudaStream_t st) {
    using G = GridBulk<D, FV, STAGES_, TPB>;
    ENSURE_DYN_SMEM((k_fold_grid_bulk<D, FV, STAGES_, TPB>), G::SMEM);
    const uint64_t mq = m / 8, n_tiles = mq / TPB;
    const int grid = (int)(n_tiles < (uint64_t)sm ? n_tiles : (uint64_t)sm);
    const FoldArgs fa = fold_args<FV>(r1, r2);
    k_fold_grid_bulk<D, FV, STAGES_, TPB><<<grid, G::THREADS_ALL, G::SMEM, st>>>(ps, mq, fa.a[0], fa.a[1], fa.a[2], fa.a[3], mb);
}

// ring shape of the bulk version: the deepest ring that fits 227 KB with 256-unit (2 KB) slabs; 128-unit slabs when that
// would leave fewer than 2 stages... (D = 3, two folded variables: 48 slabs = 96 KB per stage, 2 stages)
template <int D, int FV>
static void fold_grid_bulk_launch(const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm, cudaStream_t st) {
    constexpr int NSLOT = D * 4 * (1 << FV);
    static const int cfg = tune("ZB_GRID_BULK_TPB", 256);
    if (cfg == 128) {
        constexpr int ST = (220 * 1024) / (NSLOT * 128 * 8) > 8 ? 8 : (220 * 1024) / (NSLOT * 128 * 8);
        return fold_grid_bulk_launch_s<D, FV, ST, 128>(ps, m, r1, r2, mb, sm, st);
    }
    constexpr int ST = (220 * 1024) / (NSLOT * 256 * 8) > 6 ? 6 : (220 * 1024) / (NSLOT * 256 * 8);
    return fold_grid_bulk_launch_s<D, FV, ST, 256>(ps, m, r1, r2, mb, sm, st);
}

template <int D, int FV, int STAGES_, int TPB, bool DOT = false>
static void fold_grid_async_launch_s(const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm, cudaStream_t st) {
    using G = GridAsync<D, FV, STAGES_, TPB>;
    ENSURE_DYN_SMEM((k_fold_grid_async<D, FV, STAGES_, TPB, DOT>), G::SMEM);
    const uint64_t mq = m / 8;
    int per_sm = (227 * 1024) / (G::SMEM + 2048);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    uint64_t need = (mq + TPB - 1) / TPB, cap = (uint64_t)sm * per_sm;
    const int grid = (int)(need < cap ? (need ? need : 1) : cap);
    FoldArgs fa = fold_args<FV>(r1, r2);
    if (DOT) {
        fa.a[0] = bb::mul(bb::sub(1u, r1), bb::R_MOD_P);
        fa.a[1] = bb::mul(r1, bb::R_MOD_P);
    }
    k_fold_grid_async<D, FV, STAGES_, TPB, DOT><<<grid, TPB, G::SMEM, st>>>(ps, mq, fa.a[0], fa.a[1], fa.a[2], fa.a[3], mb);
}

// Ring shape per (D, FV): the slots of one iteration are D*4*2^FV*8 bytes per thread. Measured for d = 3
// (profiles/r01_grid_sweep.txt, tools/sweep3.sh): one folded variable: 384 threads x 2 stages 6488 GB/s, 512 x 2 6286,
// 128 x 4 6261, 256 x 3 6118, 256 x 4 5925, 384 x 3 4754; two folded variables: 256 x 2 5648, 128 x 4 5369, 192 x 2 4787.
// ZB_GRID_CFG_F1=803 selects 256 x 3 for one folded variable.
template <int D, int FV>
static void fold_grid_async_launch(const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm, cudaStream_t st) {
    constexpr int SLOT_BYTES = D * 4 * (1 << FV) * 8; // per thread and stage
    if constexpr (FV == 1) {
        static const int cfg = tune("ZB_GRID_CFG_F1", 1202);
        static const int dot = tune("ZB_GRID_DOT2", 0);
        if (cfg == 803) return fold_grid_async_launch_s<D, FV, 3, 256>(ps, m, r1, r2, mb, sm, st);
        if (dot) return fold_grid_async_launch_s<D, FV, 2, 384, true>(ps, m, r1, r2, mb, sm, st);
        return fold_grid_async_launch_s<D, FV, 2, 384>(ps, m, r1, r2, mb, sm, st);
    } else if constexpr (SLOT_BYTES * 256 * 3 <= 222 * 1024) {
        return fold_grid_async_launch_s<D, FV, 3, 256>(ps, m, r1, r2, mb, sm, st);
    } else {
        return fold_grid_async_launch_s<D, FV, 2, 256>(ps, m, r1, r2, mb, sm, st);
    }
}

template <int D, int VEC>
static void fold_grid_v(int nfold, const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm, cudaStream_t st) {
    static const int CPS = tune("ZB_GRID_CPS", 4);
    const uint64_t mq = m / (4 * VEC);
    const int grid = grid_for(mq, sm, CPS);
    const FoldArgs f1 = fold_args<1>(r1, r2), f2 = fold_args<2>(r1, r2);
    if (nfold == 0) k_fold_grid<D, 0, VEC><<<grid, THREADS, 0, st>>>(ps, mq, f1.a[0], f1.a[1], f1.a[2], f1.a[3], mb);
    else if (nfold == 1) k_fold_grid<D, 1, VEC><<<grid, THREADS, 0, st>>>(ps, mq, f1.a[0], f1.a[1], f1.a[2], f1.a[3], mb);
    else k_fold_grid<D, 2, VEC><<<grid, THREADS, 0, st>>>(ps, mq, f2.a[0], f2.a[1], f2.a[2], f2.a[3], mb);
}

template <int D>
static void fold_grid_t(int nfold, const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm, cudaStream_t st) {
    static const int ASYNC = tune("ZB_GRID_ASYNC", 1);
    static const int ASYNC_MIN = tune("ZB_GRID_ASYNC_MIN_LOG2", 16);
    // bulk-copy ring (cp.async.bulk + mbarrier): default for two folded variables, where the per-thread LDGSTS ring spends
    // 48 copy instructions per thread and iteration (ZB_GRID_BULK: bit 0 = nfold 0, bit 1 = nfold 1, bit 2 = nfold 2).
    // Measured (profiles/r02_sweep1.txt, r02_sweep5.txt, r02_ncu_full_2p30_summary.txt): at 2^29 entries the bulk ring needs
    // 13 % fewer warp-instructions and runs at 6.58 TB/s under ncu (LDGSTS ring 6.43), but in the bench loop the whole chain is
    // within 1 % either way (DRAM-bound), and for tiles counts below ~50 per CTA its 96 KB stages fill and drain too slowly:
    // bulk from 2^26 folded entries up, the per-thread ring below.
    static const int BULK = tune("ZB_GRID_BULK", 4);
    static const int BULK_MIN = tune("ZB_GRID_BULK_MIN_LOG2", 26);
    if (((BULK >> nfold) & 1) && m >= (1ull << BULK_MIN) && (m / 8) % 256 == 0) {
        if (nfold == 0) fold_grid_bulk_launch<D, 0>(ps, m, r1, r2, mb, sm, st);
        else if (nfold == 1) fold_grid_bulk_launch<D, 1>(ps, m, r1, r2, mb, sm, st);
        else fold_grid_bulk_launch<D, 2>(ps, m, r1, r2, mb, sm, st);
        return;
    }
    if (ASYNC && m >= (1ull << ASYNC_MIN)) { // the ring pays off once the pass is bandwidth-bound
        if (nfold == 0) fold_grid_async_launch<D, 0>(ps, m, r1, r2, mb, sm, st);
        else if (nfold == 1) fold_grid_async_launch<D, 1>(ps, m, r1, r2, mb, sm, st);
        else fold_grid_async_launch<D, 2>(ps, m, r1, r2, mb, sm, st);
        return;
    }
    fold_grid_v<D, 2>(nfold, ps, m, r1, r2, mb, sm, st);
}

void launch_fold_grid(int d, int nfold, const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm,
                      cudaStream_t st) {
    if (d == 1) fold_grid_t<1>(nfold, ps, m, r1, r2, mb, sm, st);
    else if (d == 2) fold_grid_t<2>(nfold, ps, m, r1, r2, mb, sm, st);
    else fold_grid_t<3>(nfold, ps, m, r1, r2, mb, sm, st);
}

// ---------------------------------------------------------------------------------------------
// Persistent tail: once the tables are small (n <= 2^14 by default) a launch per round is pure latency
// (launch + drain ~ 10 us against < 2 us of work). One CTA stays resident for ALL remaining rounds: thread 0 polls a
// host-mapped word for the next challenge (tag = expected sequence number, so a stale value can never match), the CTA
// folds in place (global memory, L2-resident) and publishes the next round's sums through the same mailbox. The host
// side keeps the transcript and only swaps "launch + wait" for "store challenge + wait".
// The kernel leaves on its own after the last round, on an abort tag, or after ~2 s without a challenge.
// ---------------------------------------------------------------------------------------------
constexpr int TAIL_THREADS = 1024;

template <int D>
__global__ void __launch_bounds__(TAIL_THREADS) k_tail_rounds(PolySet ps, uint64_t n, Mailbox mb, const unsigned long long *chal,
                                                              unsigned int chal_seq0, unsigned int *status) {
    constexpr int NS = NSums<D>::value;
    __shared__ unsigned long long sm[NS][TAIL_THREADS / 32];
    __shared__ uint32_t s_r, s_rp;
    __shared__ int s_abort;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    volatile unsigned long long *mail = (volatile unsigned long long *)mb.mail;
    for (unsigned int round = 0; n >= 2; round++, n >>= 1) {
        if (tid == 0) {
            const unsigned int want = chal_seq0 + round;
            const long long t0 = clock64();
            int ab = 0;
            unsigned long long w;
            for (;;) {
                w = *(const volatile unsigned long long *)chal;
                const unsigned int tag = (unsigned int)(w >> 32);
                if (tag == want) break;
                if (tag == TAIL_ABORT_TAG) {
                    ab = 1;
                    break;
                }
                // starvation exit: ~0.2 s for the first challenge (it is written right after the launch returns; a
                // profiler that serialises launches never lets the host get there), ~2 s for the later ones
                if (clock64() - t0 > (round == 0 ? 400000000ll : 4000000000ll)) {
                    ab = 2;
                    break;
                }
            }
            s_abort = ab;
            s_r = (uint32_t)w;
            s_rp = bb::shoup_pre((uint32_t)w);
        }
        __syncthreads();
        if (s_abort) {
            if (tid == 0 && s_abort == 2) *status = 1u;
            return;
        }
        const uint32_t r = s_r, rp = s_rp;
        if (n == 2) { // last fold: one value per polynomial remains; payload = the D final evaluations
            if (tid == 0) {
#pragma unroll
                for (int k = 0; k < D; k++) {
                    uint32_t v = bb::lerp(ps.src[k][0], ps.src[k][1], r, rp);
                    ps.dst[k][0] = v;
                    mail[k] = mb.tagged ? mail_tagged(mb.seq + round, v) : v;
                }
                if (!mb.tagged) __threadfence_system();
                mail[MAIL_WORDS] = mb.seq + round;
            }
            return;
        }
        unsigned long long s[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) s[k] = 0;
        const uint64_t q = n / 4;
        if ((q & 3) == 0) {
            const uint64_t q4 = q / 4;
            for (uint64_t i = tid; i < q4; i += TAIL_THREADS) {
                uint32_t nlo[D][4], nhi[D][4];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const uint4 *p = reinterpret_cast<const uint4 *>(ps.src[k]);
                    uint4 v0 = p[i], v1 = p[i + q4], v2 = p[i + 2 * q4], v3 = p[i + 3 * q4];
                    uint32_t e0[4], e1[4], e2[4], e3[4];
                    UNPACK4(v0, e0);
                    UNPACK4(v1, e1);
                    UNPACK4(v2, e2);
                    UNPACK4(v3, e3);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        nlo[k][c] = bb::lerp(e0[c], e2[c], r, rp);
                        nhi[k][c] = bb::lerp(e1[c], e3[c], r, rp);
                    }
                    uint4 *o = reinterpret_cast<uint4 *>(ps.dst[k]);
                    o[i] = make_uint4(nlo[k][0], nlo[k][1], nlo[k][2], nlo[k][3]);
                    o[i + q4] = make_uint4(nhi[k][0], nhi[k][1], nhi[k][2], nhi[k][3]);
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t l[D], h[D];
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        l[k] = nlo[k][c];
                        h[k] = nhi[k][c];
                    }
                    accum_pair<D>(l, h, s);
                }
            }
        } else {
            for (uint64_t i = tid; i < q; i += TAIL_THREADS) {
                uint32_t l[D], h[D];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const uint32_t *p = ps.src[k];
                    uint32_t e0 = p[i], e1 = p[i + q], e2 = p[i + 2 * q], e3 = p[i + 3 * q];
                    l[k] = bb::lerp(e0, e2, r, rp);
                    h[k] = bb::lerp(e1, e3, r, rp);
                    ps.dst[k][i] = l[k];
                    ps.dst[k][i + q] = h[k];
                }
                accum_pair<D>(l, h, s);
            }
        }
#pragma unroll
        for (int k = 0; k < NS; k++) {
            unsigned long long v = warp_sum(s[k]);
            if (lane == 0) sm[k][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
            unsigned long long tot[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) tot[k] = warp_sum(sm[k][lane]);
            if (lane == 0) {
                Finish<D>()(tot);
#pragma unroll
                for (int k = 0; k < NS; k++) mail[k] = mb.tagged ? mail_tagged(mb.seq + round, tot[k]) : tot[k];
                if (!mb.tagged) __threadfence_system();
                mail[MAIL_WORDS] = mb.seq + round;
            }
        }
        __syncthreads(); // this round's stores are visible to the whole CTA before the next round reads them
    }
}

// Small tables leave the device: bind `nfold` (0..2) top variables of the D tables (n = m << nfold entries each, the same dot
// products as the grid kernels) and write the m folded values of every table, canonical u32, table after table, to `dump`
// (host-mapped) — the host twin finishes the last <= 10 rounds of a product sumcheck itself instead of paying a host round
// trip (~8 us) per round for a few hundred multiplications. One CTA: m <= 2^PROD_DUMP_MAX_LOG2. The device tables stay as they are.
template <int D>
__global__ void __launch_bounds__(TAIL_THREADS) k_fold_dump(PolySet ps, uint64_t m, int nfold, bb::BilinearWeights bw, uint32_t *dump,
                                                            Mailbox mb) {
    const uint32_t w0 = bw.w[0], w1 = bw.w[1], w2 = bw.w[2], w3 = bw.w[3];
    for (uint64_t i = threadIdx.x; i < m; i += TAIL_THREADS) {
#pragma unroll
        for (int k = 0; k < D; k++) {
            const uint32_t *p = ps.src[k];
            uint32_t v;
            if (nfold == 0) v = p[i];
            else if (nfold == 1) v = bb::dot2(p[i], p[i + m], w0, w1);
            else v = bb::dot4(p[i], p[i + m], p[i + 2 * m], p[i + 3 * m], w0, w1, w2, w3);
            asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(dump + (uint64_t)k * m + i), "r"(v) : "memory");
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) ((volatile unsigned long long *)mb.mail)[MAIL_WORDS] = mb.seq;
}

void launch_fold_dump(int d, int nfold, const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, uint32_t *dump, const Mailbox &mb,
                      cudaStream_t st) {
    bb::BilinearWeights bw{};
    if (nfold == 1) {
        bw.w[0] = bb::mul(bb::sub(1u, r1), bb::R_MOD_P);
        bw.w[1] = bb::mul(r1, bb::R_MOD_P);
    } else if (nfold == 2) {
        bw = bb::bilinear_weights(r1, r2);
    }
    if (d == 1) k_fold_dump<1><<<1, TAIL_THREADS, 0, st>>>(ps, m, nfold, bw, dump, mb);
    else if (d == 2) k_fold_dump<2><<<1, TAIL_THREADS, 0, st>>>(ps, m, nfold, bw, dump, mb);
    else k_fold_dump<3><<<1, TAIL_THREADS, 0, st>>>(ps, m, nfold, bw, dump, mb);
}

// the tables of a consumed product prove end as their final evaluations (what v in-place folds would have left in slot 0)
__global__ void k_fill_heads(PolySet ps, int d, uint32_t v0, uint32_t v1, uint32_t v2) {
    if (threadIdx.x < d) ps.dst[threadIdx.x][0] = threadIdx.x == 0 ? v0 : (threadIdx.x == 1 ? v1 : v2);
}
void launch_fill_heads(const PolySet &ps, int d, const uint32_t *vals, cudaStream_t st) {
    k_fill_heads<<<1, 32, 0, st>>>(ps, d, vals[0], d > 1 ? vals[1] : 0, d > 2 ? vals[2] : 0);
}

void launch_tail_rounds(int d, const PolySet &ps, uint64_t n, const Mailbox &mb, const unsigned long long *chal,
                        unsigned int chal_seq0, unsigned int *status, cudaStream_t st) {
    if (d == 1) k_tail_rounds<1><<<1, TAIL_THREADS, 0, st>>>(ps, n, mb, chal, chal_seq0, status);
    else if (d == 2) k_tail_rounds<2><<<1, TAIL_THREADS, 0, st>>>(ps, n, mb, chal, chal_seq0, status);
    else k_tail_rounds<3><<<1, TAIL_THREADS, 0, st>>>(ps, n, mb, chal, chal_seq0, status);
}

// multi-GPU helpers: publish NCCL-summed payload words (canonical values summed over <= 16 ranks: < 2^35) mod p
__global__ void k_publish_reduced(const unsigned long long *src, int n, unsigned long long *mail, unsigned long long seq) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int k = 0; k < n; k++) ((volatile unsigned long long *)mail)[k] = src[k] % bb::P;
        __threadfence_system();
        ((volatile unsigned long long *)mail)[MAIL_WORDS] = seq;
    }
}
void launch_publish_reduced(const unsigned long long *src, int n, unsigned long long *mail, unsigned long long seq, cudaStream_t st) {
    k_publish_reduced<<<1, 32, 0, st>>>(src, n, mail, seq);
}

// all-gathered cyclic shards [rank][j] -> global order out[rank + world * j]
__global__ void k_interleave(const uint32_t *gathered, uint32_t *out, uint64_t n_local, uint32_t world) {
    const uint64_t total = n_local * world;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        out[i] = gathered[(i % world) * n_local + i / world];
}
void launch_interleave(const uint32_t *gathered, uint32_t *out, uint64_t n_local, uint32_t world, cudaStream_t st) {
    uint64_t g = (n_local * world + 255) / 256;
    k_interleave<<<(int)(g > 148 * 16 ? 148 * 16 : (g ? g : 1)), 256, 0, st>>>(gathered, out, n_local, world);
}


// ---- single-process multi-GPU (zb_ctx_create_mask): tables move between the GPUs' layouts through peer memory ----
// pull: out[j * world + q] = src_q[j] (j < n_local): every warp reads 128 contiguous bytes from each peer over NVLink,
// the interleave happens in shared memory, the local writes are fully coalesced. Serves the all-gather of cyclic shards
// and the cyclic -> contiguous-block transpose in front of a sharded Merkle build.
constexpr int PEER_J = 256;
__global__ void __launch_bounds__(PEER_J) k_interleave_peers(PeerSrc ps, uint32_t *out, uint64_t n_local, uint32_t world) {
    extern __shared__ uint32_t sm_peer[]; // [PEER_J * world]
    for (uint64_t j0 = (uint64_t)blockIdx.x * PEER_J; j0 < n_local; j0 += (uint64_t)gridDim.x * PEER_J) {
        const uint64_t j = j0 + threadIdx.x;
        for (uint32_t q = 0; q < world; q++) sm_peer[threadIdx.x * world + q] = j < n_local ? ps.p[q][j] : 0u;
        __syncthreads();
        const uint64_t cnt = (n_local - j0 < PEER_J ? n_local - j0 : PEER_J) * world;
        for (uint64_t k = threadIdx.x; k < cnt; k += PEER_J) out[j0 * world + k] = sm_peer[k];
        __syncthreads();
    }
}
void launch_interleave_peers(const PeerSrc &ps, uint32_t *out, uint64_t n_local, uint32_t world, int sm, cudaStream_t st) {
    uint64_t g = (n_local + PEER_J - 1) / PEER_J, cap = (uint64_t)sm * 8;
    k_interleave_peers<<<(int)(g < cap ? (g ? g : 1) : cap), PEER_J, PEER_J * world * sizeof(uint32_t), st>>>(ps, out, n_local, world);
}
// The all-gather that ends the sharded regime, without NCCL, host rendezvous or stream synchronisation (launch_gather_xchg in
// kernels.h). Phase 1: the grid copies this rank's shards into its own staging set; the last CTA to finish (atomic ticket)
// raises this rank's arrival word in every peer's buffer (after a system-scope fence). Phase 2: every CTA waits for all
// arrival words in the LOCAL buffer, then pulls the peers' staged shards — coalesced remote reads, interleave in shared memory,
// coalesced local writes, as k_interleave_peers. The grid is small enough to be co-resident (the wait cannot starve phase 1).
constexpr int GATHER_CTAS = 148; // one per SM at most: the whole grid must be resident while it waits for the peers
__device__ __forceinline__ uint4 ld_volatile_v4(const uint32_t *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// pull for shard lengths that are multiples of 4: one 16-byte load per peer (all W of them in flight), then the W x 4 values a
// thread holds are 4 W CONSECUTIVE output words (rows j .. j + 3 of W entries each): 16-byte stores, no shared memory
template <int W>
__device__ __forceinline__ void gather_pull_v4(const XchgView *xv, size_t set, int count, const PolySet &outs, uint64_t n4) {
    const uint64_t items = (uint64_t)count * n4, stride = (uint64_t)gridDim.x * PEER_J;
    for (uint64_t it = (uint64_t)blockIdx.x * PEER_J + threadIdx.x; it < items; it += stride) {
        const int k = (int)(it / n4);
        const uint64_t j4 = it - (uint64_t)k * n4;
        uint32_t v[W][4];
#pragma unroll
        for (int q = 0; q < W; q++) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(xv->peer[q] + XCHG_GATHER_DATA_OFF) + set * XCHG_GATHER_SET_ELEMS +
                                  ((size_t)k << XCHG_GATHER_MAX_LOG2);
            const uint4 x = ld_volatile_v4(src + 4 * j4);
            v[q][0] = x.x, v[q][1] = x.y, v[q][2] = x.z, v[q][3] = x.w;
        }
        uint4 *o = reinterpret_cast<uint4 *>(outs.dst[k] + 4 * j4 * W);
        if constexpr (W == 2) {
            o[0] = make_uint4(v[0][0], v[1][0], v[0][1], v[1][1]);
            o[1] = make_uint4(v[0][2], v[1][2], v[0][3], v[1][3]);
        } else {
#pragma unroll
            for (int jj = 0; jj < 4; jj++)
#pragma unroll
                for (int q = 0; q < W; q += 4) o[(jj * W + q) / 4] = make_uint4(v[q][jj], v[q + 1][jj], v[q + 2][jj], v[q + 3][jj]);
        }
    }
}

__global__ void __launch_bounds__(PEER_J) k_gather_xchg(const XchgView *xv, PolySet shards, PolySet outs, int count, uint64_t n_local,
                                                        unsigned long long gseq, unsigned int *ticket, unsigned long long *mail) {
    extern __shared__ uint32_t sm_peer[]; // [PEER_J * world] (scalar path only)
    __shared__ int s_ok;
    const int world = xv->world, rank = xv->rank;
    const size_t set = (size_t)(gseq & 1ull);
    unsigned long long *mine = xv->peer[rank];
    uint32_t *stage = reinterpret_cast<uint32_t *>(mine + XCHG_GATHER_DATA_OFF) + set * XCHG_GATHER_SET_ELEMS;
    const uint64_t gtid = (uint64_t)blockIdx.x * PEER_J + threadIdx.x, gstride = (uint64_t)gridDim.x * PEER_J;
    const bool vec = (n_local % 4) == 0 && (world == 2 || world == 4 || world == 8 || world == 16);
    if (vec) {
        const uint64_t n4 = n_local / 4;
        for (uint64_t it = gtid; it < (uint64_t)count * n4; it += gstride) {
            const int k = (int)(it / n4);
            const uint64_t j4 = it - (uint64_t)k * n4;
            reinterpret_cast<uint4 *>(stage + ((size_t)k << XCHG_GATHER_MAX_LOG2))[j4] = reinterpret_cast<const uint4 *>(shards.src[k])[j4];
        }
    } else {
        for (int k = 0; k < count; k++)
            for (uint64_t i = gtid; i < n_local; i += gstride) stage[((size_t)k << XCHG_GATHER_MAX_LOG2) + i] = shards.src[k][i];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) { // every CTA's copies are in place
            *ticket = 0u;
            __threadfence_system();
            for (int q = 0; q < world; q++)
                ((volatile unsigned long long *)xv->peer[q])[XCHG_GATHER_FLAGS_OFF + set * XCHG_MAX_RANKS + rank] = gseq;
        }
    }
    if (threadIdx.x < 32) {
        bool ok = true;
        if ((int)threadIdx.x < world) {
            volatile unsigned long long *flag = mine + XCHG_GATHER_FLAGS_OFF + set * XCHG_MAX_RANKS + threadIdx.x;
            const long long t0 = clock64(), patience = xv->patience;
            while (*flag != gseq)
                if (clock64() - t0 > patience) {
                    ok = false;
                    break;
                }
        }
        ok = __all_sync(0xffffffffu, ok);
        if (threadIdx.x == 0) {
            s_ok = ok ? 1 : 0;
            if (!ok) { // reported by the next wait on the mailbox ("a rank never arrived")
                ((volatile unsigned long long *)mail)[MAIL_WORDS - 1] = 1ull;
                __threadfence_system();
            }
        }
    }
    __syncthreads();
    if (!s_ok) return;
    __threadfence_system();
    if (vec) {
        const uint64_t n4 = n_local / 4;
        if (world == 2) gather_pull_v4<2>(xv, set, count, outs, n4);
        else if (world == 4) gather_pull_v4<4>(xv, set, count, outs, n4);
        else if (world == 8) gather_pull_v4<8>(xv, set, count, outs, n4);
        else gather_pull_v4<16>(xv, set, count, outs, n4);
        return;
    }
    for (int k = 0; k < count; k++) {
        uint32_t *out = outs.dst[k];
        for (uint64_t j0 = (uint64_t)blockIdx.x * PEER_J; j0 < n_local; j0 += (uint64_t)gridDim.x * PEER_J) {
            const uint64_t j = j0 + threadIdx.x;
            for (int q = 0; q < world; q++) {
                const volatile uint32_t *src = reinterpret_cast<const volatile uint32_t *>(xv->peer[q] + XCHG_GATHER_DATA_OFF) +
                                               set * XCHG_GATHER_SET_ELEMS + ((size_t)k << XCHG_GATHER_MAX_LOG2);
                sm_peer[threadIdx.x * world + q] = j < n_local ? src[j] : 0u;
            }
            __syncthreads();
            const uint64_t cnt = (n_local - j0 < PEER_J ? n_local - j0 : PEER_J) * world;
            for (uint64_t t = threadIdx.x; t < cnt; t += PEER_J) out[j0 * world + t] = sm_peer[t];
            __syncthreads();
        }
    }
}
void launch_gather_xchg(const XchgView *xv, int world, const PolySet &shards, const PolySet &outs, int count, uint64_t n_local,
                        unsigned long long gseq, unsigned int *ticket, unsigned long long *mail, cudaStream_t st) {
    // vector path: count * n_local / 4 items; scalar path: tiles of PEER_J entries
    const uint64_t g = (n_local % 4) == 0 ? ((uint64_t)count * (n_local / 4) + PEER_J - 1) / PEER_J : (n_local + PEER_J - 1) / PEER_J;
    const int grid = (int)(g < (uint64_t)GATHER_CTAS ? (g ? g : 1) : (uint64_t)GATHER_CTAS);
    k_gather_xchg<<<grid, PEER_J, PEER_J * world * sizeof(uint32_t), st>>>(xv, shards, outs, count, n_local, gseq, ticket, mail);
}
// push: dst_q[j] = src[j * world + q] (j < n_out): coalesced local reads, 128-byte coalesced stores into every peer.
// Serves the host-order (contiguous block per GPU) -> cyclic-shard deal after an upload.
__global__ void __launch_bounds__(PEER_J) k_deal_peers(const uint32_t *src, PeerDst pd, uint64_t n_out, uint32_t world) {
    extern __shared__ uint32_t sm_peer[];
    for (uint64_t j0 = (uint64_t)blockIdx.x * PEER_J; j0 < n_out; j0 += (uint64_t)gridDim.x * PEER_J) {
        const uint64_t cnt = (n_out - j0 < PEER_J ? n_out - j0 : PEER_J) * world;
        for (uint64_t k = threadIdx.x; k < cnt; k += PEER_J) sm_peer[k] = src[j0 * world + k];
        __syncthreads();
        const uint64_t j = j0 + threadIdx.x;
        if (j < n_out)
            for (uint32_t q = 0; q < world; q++) pd.p[q][j] = sm_peer[threadIdx.x * world + q];
        __syncthreads();
    }
}
void launch_deal_peers(const uint32_t *src, const PeerDst &pd, uint64_t n_out, uint32_t world, int sm, cudaStream_t st) {
    uint64_t g = (n_out + PEER_J - 1) / PEER_J, cap = (uint64_t)sm * 8;
    k_deal_peers<<<(int)(g < cap ? (g ? g : 1) : cap), PEER_J, PEER_J * world * sizeof(uint32_t), st>>>(src, pd, n_out, world);
}

template <int D>
static void round_sums_t(const PolySet &ps, uint64_t n, const Mailbox &mb, int sm, cudaStream_t st) {
    uint64_t h = n / 2;
    if (h % 4 == 0) {
        static const int CPS = tune("ZB_RSUM_CPS", 16);
        k_round_sums_v4<D><<<grid_for(h / 4, sm, CPS), THREADS, 0, st>>>(ps, h / 4, mb);
    } else {
        k_round_sums_s<D><<<1, THREADS, 0, st>>>(ps, h, mb);
    }
}

void launch_round_sums(int d, const PolySet &ps, uint64_t n, const Mailbox &mb, int sm, cudaStream_t st) {
    if (d == 1) round_sums_t<1>(ps, n, mb, sm, st);
    else if (d == 2) round_sums_t<2>(ps, n, mb, sm, st);
    else round_sums_t<3>(ps, n, mb, sm, st);
}

template <int D>
static void fold_sums_t(const PolySet &ps, uint64_t n, uint32_t r, const Mailbox &mb, int sm, cudaStream_t st, const ChalSrc *csp) {
    uint32_t rp = bb::shoup_pre(r);
    const ChalSrc cs = csp ? *csp : ChalSrc{nullptr, nullptr, nullptr, 0};
    if (n == 2) {
        k_fold_last<D><<<1, 32, 0, st>>>(ps, r, rp, mb);
        return;
    }
    uint64_t q = n / 4;
    if (q % 4 == 0) {
        static const int U = tune("ZB_FOLD_U", D == 1 ? 4 : D == 2 ? 2 : 1);
        static const int CPS = tune("ZB_FOLD_CPS", 8);
        const uint64_t q4 = q / 4;
        if (U >= 4 && D == 1)
            k_fold_sums_v4<D, (D == 1 ? 4 : 1)><<<grid_for((q4 + 3) / 4, sm, CPS), THREADS, 0, st>>>(ps, q4, r, rp, mb, cs);
        else if (U >= 2 && D <= 2)
            k_fold_sums_v4<D, (D <= 2 ? 2 : 1)><<<grid_for((q4 + 1) / 2, sm, CPS), THREADS, 0, st>>>(ps, q4, r, rp, mb, cs);
        else
            k_fold_sums_v4<D, 1><<<grid_for(q4, sm, CPS), THREADS, 0, st>>>(ps, q4, r, rp, mb, cs);
    } else {
        k_fold_sums_s<D><<<1, THREADS, 0, st>>>(ps, q, r, rp, mb);
    }
}

void launch_fold_sums(int d, const PolySet &ps, uint64_t n, uint32_t r, const Mailbox &mb, int sm, cudaStream_t st,
                      const ChalSrc *cs) {
    if (d == 1) fold_sums_t<1>(ps, n, r, mb, sm, st, cs);
    else if (d == 2) fold_sums_t<2>(ps, n, r, mb, sm, st, cs);
    else fold_sums_t<3>(ps, n, r, mb, sm, st, cs);
}

// ---------------------------------------------------------------------------------------------
// plain sum (sumOverHypercube)
// ---------------------------------------------------------------------------------------------
struct FinishSum {
    __device__ void operator()(unsigned long long (&t)[1]) const { t[0] = bb::reduce64_scaled<0>(t[0]); }
};

__global__ void __launch_bounds__(THREADS) k_sum(const uint32_t *src, uint64_t n, Mailbox mb) {
    unsigned long long s[1] = {0};
    const uint64_t stride = (uint64_t)gridDim.x * THREADS;
    const uint64_t n4 = n / 4;
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
#pragma unroll 4
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n4; i += stride) {
        uint4 v = p[i];
        s[0] += (unsigned long long)(v.x + v.y) + (unsigned long long)(v.z + v.w); // each pair < 2^32
    }
    for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) s[0] += src[i];
    publish_sums<1>(s, mb, FinishSum());
}

void launch_sum(const uint32_t *src, uint64_t n, const Mailbox &mb, int sm, cudaStream_t st) {
    k_sum<<<grid_for((n + 3) / 4, sm, 8), THREADS, 0, st>>>(src, n, mb);
}

// ---------------------------------------------------------------------------------------------
// d = 1: several rounds per pass through linearity (see kernels.h). The CTAs sweep the table in grid-stride order (all
// CTAs next to each other in every operand stream: one DRAM/TLB hot spot per stream, like the other fold kernels — a first
// version that gave every CTA its own output block kept 32 x 32 spots open and reached 4.4-5.2 TB/s). Block sums: a
// warp's 32 consecutive vectors always lie in one block (block length >= 128 elements), so the warp adds its running sum to
// a shared-memory accumulator whenever its block changes; one global atomicAdd per CTA and block at the end.
// ---------------------------------------------------------------------------------------------
constexpr int LIN_NB = 1 << LIN_MAX_K;

struct BlockAcc {
    unsigned long long *sacc; // shared: LIN_NB accumulators
    __device__ __forceinline__ void init(unsigned long long *sm) {
        sacc = sm;
        if (threadIdx.x < LIN_NB) sacc[threadIdx.x] = 0;
        __syncthreads();
    }
    // warp-uniform block index
    __device__ __forceinline__ void flush_warp(uint64_t b, unsigned long long s) {
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&sacc[b], s);
    }
    __device__ __forceinline__ void flush_lane(uint64_t b, unsigned long long s) {
        if (s) atomicAdd(&sacc[b], s);
    }
    // all threads: CTA totals -> global accumulators; the last CTA publishes nb canonical sums (nb == 0: only the sequence number)
    // nb == 0 (a table was published instead of sums): the CTA's host-mapped stores are made visible by ONE system
    // fence of thread 0 behind the CTA barrier (fences are cumulative) instead of one per thread.
    __device__ __forceinline__ void publish(int nb, const Mailbox &mb) {
        __shared__ bool is_last;
        __syncthreads();
        if ((int)threadIdx.x < nb && sacc[threadIdx.x]) {
            atomicAdd(&mb.acc[threadIdx.x], sacc[threadIdx.x]);
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (nb == 0) __threadfence_system();
            else __threadfence();
            is_last = (atomicAdd(mb.ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last && threadIdx.x < 32) {
            __threadfence();
            const int lane = threadIdx.x;
            volatile unsigned long long *mail = (volatile unsigned long long *)mb.mail;
            if (lane < nb) {
                const unsigned long long v = bb::reduce64_scaled<0>(atomicExch(&mb.acc[lane], 0ull)); // read + re-arm
                mail[lane] = mb.tagged ? mail_tagged(mb.seq, v) : v;
            }
            if (lane == 0) *mb.ticket = 0u;
            if (!mb.tagged || nb == 0) __threadfence_system(); // plain words / a published table: payload before the sequence number
            __syncwarp();
            if (lane == 0) mail[MAIL_WORDS] = mb.seq;
        }
    }
};

constexpr int LIN_TPB = 256;

// n4 uint4 in the table, blocks of 2^logL4 uint4 each
// chunk4 == 0: grid-stride order (large tables). chunk4 > 0 (small, L2-resident tables): CTA c owns the contiguous vectors
// [c chunk4, (c+1) chunk4), so a thread's loads stay inside one block and are all in flight together.
__global__ void __launch_bounds__(LIN_TPB) k_block_sums(const uint32_t *__restrict__ src, uint64_t n4, int logL4, int nb, uint64_t chunk4,
                                                        Mailbox mb) {
    __shared__ unsigned long long sm[LIN_NB];
    BlockAcc acc;
    acc.init(sm);
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
    const uint64_t stride = chunk4 ? LIN_TPB : (uint64_t)gridDim.x * LIN_TPB;
    uint64_t i = (chunk4 ? (uint64_t)blockIdx.x * chunk4 : (uint64_t)blockIdx.x * LIN_TPB) + threadIdx.x;
    const uint64_t limit = chunk4 && (blockIdx.x + 1) * chunk4 < n4 ? (blockIdx.x + 1) * chunk4 : n4;
    while (i < limit) {
        const uint64_t b = i >> logL4;
        const uint64_t bend = (b + 1) << logL4; // <= n4
        const uint64_t end = bend < limit ? bend : limit;
        unsigned long long s = 0;
#pragma unroll 8
        for (; i < end; i += stride) {
            const uint4 v = p[i];
            s += (unsigned long long)(v.x + v.y) + (unsigned long long)(v.z + v.w); // each pair < 2^32
        }
        if (logL4 >= 5) acc.flush_warp(b, s);
        else acc.flush_lane(b, s);
    }
    acc.publish(nb, mb);
}

void launch_block_sums(const uint32_t *src, uint64_t n, int k, const Mailbox &mb, int sm, cudaStream_t st) {
    static const int CPS = tune("ZB_BSUM_CPS", 8);
    const uint64_t n4 = n / 4;
    int logL4 = 0;
    while ((1ull << (logL4 + k)) < n4) logL4++;
    // small tables: at least 8 loads per thread (all in flight together) instead of one CTA per 256 vectors — the pass is
    // latency-bound there and every CTA costs a ticket atomic
    const uint64_t cap = (uint64_t)sm * CPS, chunk4 = 8 * LIN_TPB;
    if (n4 <= cap * chunk4) {
        k_block_sums<<<(int)((n4 + chunk4 - 1) / chunk4), LIN_TPB, 0, st>>>(src, n4, logL4, 1 << k, chunk4, mb);
    } else {
        k_block_sums<<<(int)cap, LIN_TPB, 0, st>>>(src, n4, logL4, 1 << k, 0, mb);
    }
}

// sum of <= 8 canonical values -> canonical: q = x >> 31 never exceeds floor(x / P) and x - q P < 2 P
__device__ __forceinline__ uint32_t reduce_small(unsigned long long x) {
    const uint32_t q = (uint32_t)(x >> 31);
    const uint32_t r = (uint32_t)x - q * bb::P;
    return min(r, r - bb::P);
}

template <int K, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_foldk_sums(const uint32_t *src, uint32_t *dst, uint64_t m4, int logL4, int nb,
                                                          FoldWeights fw, unsigned long long *dump, Mailbox mb) {
    constexpr int NT = 1 << K;
    __shared__ unsigned long long sm[LIN_NB];
    BlockAcc acc;
    acc.init(sm);
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
    uint4 *o = reinterpret_cast<uint4 *>(dst);
    const uint64_t stride = (uint64_t)gridDim.x * TPB;
    uint64_t i4 = (uint64_t)blockIdx.x * TPB + threadIdx.x;
    while (i4 < m4) {
        const uint64_t b = dump ? 0 : i4 >> logL4;
        const uint64_t end = dump ? m4 : (b + 1) << logL4;
        unsigned long long s = 0;
        for (; i4 < end; i4 += stride) {
            uint4 v[NT];
#pragma unroll
            for (int t = 0; t < NT; t++) v[t] = p[(uint64_t)t * m4 + i4];
            uint32_t r[4];
            if constexpr (K == 1) {
                r[0] = bb::dot2(v[0].x, v[1].x, fw.w[0], fw.w[1]);
                r[1] = bb::dot2(v[0].y, v[1].y, fw.w[0], fw.w[1]);
                r[2] = bb::dot2(v[0].z, v[1].z, fw.w[0], fw.w[1]);
                r[3] = bb::dot2(v[0].w, v[1].w, fw.w[0], fw.w[1]);
            } else {
                unsigned long long a[4] = {0, 0, 0, 0};
#pragma unroll
                for (int g = 0; g < NT / 4; g++) {
                    const uint32_t w0 = fw.w[4 * g], w1 = fw.w[4 * g + 1], w2 = fw.w[4 * g + 2], w3 = fw.w[4 * g + 3];
                    a[0] += bb::dot4(v[4 * g].x, v[4 * g + 1].x, v[4 * g + 2].x, v[4 * g + 3].x, w0, w1, w2, w3);
                    a[1] += bb::dot4(v[4 * g].y, v[4 * g + 1].y, v[4 * g + 2].y, v[4 * g + 3].y, w0, w1, w2, w3);
                    a[2] += bb::dot4(v[4 * g].z, v[4 * g + 1].z, v[4 * g + 2].z, v[4 * g + 3].z, w0, w1, w2, w3);
                    a[3] += bb::dot4(v[4 * g].w, v[4 * g + 1].w, v[4 * g + 2].w, v[4 * g + 3].w, w0, w1, w2, w3);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) r[q] = NT == 4 ? (uint32_t)a[q] : reduce_small(a[q]);
            }
            o[i4] = make_uint4(r[0], r[1], r[2], r[3]);
            if (dump != nullptr) { // published table: canonical u32, one 16-byte store per thread (a quarter of the PCIe writes of u64)
                asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<uint4 *>(dump) + i4), "r"(r[0]), "r"(r[1]),
                             "r"(r[2]), "r"(r[3])
                             : "memory");
            } else {
                s += (unsigned long long)(r[0] + r[1]) + (unsigned long long)(r[2] + r[3]);
            }
        }
        if (dump == nullptr) {
            if (logL4 >= 5) acc.flush_warp(b, s);
            else acc.flush_lane(b, s);
        }
    }
    acc.publish(nb, mb);
}

template <int K>
static void foldk_launch(const uint32_t *src, uint32_t *dst, uint64_t n, const FoldWeights &w, int k_next, unsigned long long *dump,
                         const Mailbox &mb, int sm, cudaStream_t st) {
    // K = 5 keeps 32 128-bit loads per thread in flight (128 data registers): 128 threads x 3 CTAs per SM
    constexpr int TPB = K >= 5 ? 128 : 256, MINB = K >= 5 ? 3 : 2;
    static const int CPS = tune("ZB_FOLDK_CPS", 2); // measured at 2^28 (r02_sweep2.txt): 2 -> 5228 GB/s, 3 -> 5082, 4 -> 4791
    const uint64_t m4 = (n >> K) / 4;
    const int nb = dump ? 0 : (1 << k_next);
    int logL4 = 0;
    if (!dump)
        while ((1ull << (logL4 + k_next)) < m4) logL4++;
    uint64_t need = (m4 + TPB - 1) / TPB, cap = (uint64_t)sm * CPS;
    const int grid = (int)(need < cap ? need : cap);
    k_foldk_sums<K, TPB, MINB><<<grid, TPB, 0, st>>>(src, dst, m4, logL4, nb, w, dump, mb);
}

void launch_foldk_sums(const uint32_t *src, uint32_t *dst, uint64_t n, int k, const FoldWeights &w, int k_next,
                       unsigned long long *dump, const Mailbox &mb, int sm, cudaStream_t st) {
    switch (k) {
    case 1: foldk_launch<1>(src, dst, n, w, k_next, dump, mb, sm, st); break;
    case 2: foldk_launch<2>(src, dst, n, w, k_next, dump, mb, sm, st); break;
    case 3: foldk_launch<3>(src, dst, n, w, k_next, dump, mb, sm, st); break;
    case 4: foldk_launch<4>(src, dst, n, w, k_next, dump, mb, sm, st); break;
    default: foldk_launch<5>(src, dst, n, w, k_next, dump, mb, sm, st); break;
    }
}

// ---- the same two steps for small (L2-resident, latency-bound) tables: up to 8 variables per pass ----
// Block sums: CTA b owns block b; its canonical sum goes out as ONE self-validating word. No accumulators, no ticket.
__global__ void __launch_bounds__(LIN_TPB) k_block_sums_wide(const uint32_t *__restrict__ src, uint64_t len4, unsigned long long *words,
                                                             unsigned long long seq) {
    __shared__ unsigned long long sm[LIN_TPB / 32];
    const uint4 *p = reinterpret_cast<const uint4 *>(src) + (uint64_t)blockIdx.x * len4;
    unsigned long long s = 0;
#pragma unroll 4
    for (uint64_t i = threadIdx.x; i < len4; i += LIN_TPB) {
        const uint4 v = p[i];
        s += (unsigned long long)(v.x + v.y) + (unsigned long long)(v.z + v.w); // each pair < 2^32
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < LIN_TPB / 32; w++) t += sm[w];
        ((volatile unsigned long long *)words)[blockIdx.x] = mail_tagged(seq, bb::reduce64_scaled<0>(t));
    }
}
void launch_block_sums_wide(const uint32_t *src, uint64_t n, int k, unsigned long long *words, unsigned long long seq, cudaStream_t st) {
    k_block_sums_wide<<<1 << k, LIN_TPB, 0, st>>>(src, (n >> k) / 4, words, seq);
}

// Fold: a CTA owns 16 vector columns (64 outputs) and ALL 2^K rows, split over G row groups of RPG rows (RPG loads per thread,
// all in flight; dot products in groups of four, bb::dot4); the groups meet in shared memory. In place is safe: a column is
// read and written by one CTA only, and the writes follow the barrier.
template <int K>
__global__ void __launch_bounds__(LIN_TPB) k_foldk_wide(const uint32_t *src, uint32_t *dst, uint64_t m4, WideWeights ww,
                                                        unsigned long long *words, unsigned long long seq) {
    constexpr int NB = 1 << K;
    constexpr int RPG = NB >= 64 ? NB / 16 : (NB >= 4 ? 4 : NB); // rows per group: 2, 4, 4, 4, 4, 4, 8, 16
    constexpr int G = NB / RPG;                                   // groups:         1, 1, 2, 4, 8, 16, 16, 16
    __shared__ unsigned long long part[G][16][4];
    const int o = threadIdx.x & 15, g = threadIdx.x >> 4;
    const uint64_t col = (uint64_t)blockIdx.x * 16 + o;
    if (g < G) {
        unsigned long long a[4] = {0, 0, 0, 0};
        if (col < m4) {
            const uint4 *p = reinterpret_cast<const uint4 *>(src);
            uint4 v[RPG];
#pragma unroll
            for (int r = 0; r < RPG; r++) v[r] = p[(uint64_t)(g * RPG + r) * m4 + col];
            if constexpr (RPG == 2) {
                const uint32_t w0 = ww.w[0], w1 = ww.w[1];
                a[0] = bb::dot2(v[0].x, v[1].x, w0, w1);
                a[1] = bb::dot2(v[0].y, v[1].y, w0, w1);
                a[2] = bb::dot2(v[0].z, v[1].z, w0, w1);
                a[3] = bb::dot2(v[0].w, v[1].w, w0, w1);
            } else {
#pragma unroll
                for (int r = 0; r < RPG; r += 4) {
                    const uint32_t w0 = ww.w[g * RPG + r], w1 = ww.w[g * RPG + r + 1], w2 = ww.w[g * RPG + r + 2], w3 = ww.w[g * RPG + r + 3];
                    a[0] += bb::dot4(v[r].x, v[r + 1].x, v[r + 2].x, v[r + 3].x, w0, w1, w2, w3);
                    a[1] += bb::dot4(v[r].y, v[r + 1].y, v[r + 2].y, v[r + 3].y, w0, w1, w2, w3);
                    a[2] += bb::dot4(v[r].z, v[r + 1].z, v[r + 2].z, v[r + 3].z, w0, w1, w2, w3);
                    a[3] += bb::dot4(v[r].w, v[r + 1].w, v[r + 2].w, v[r + 3].w, w0, w1, w2, w3);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) part[g][o][c] = a[c];
    }
    __syncthreads();
    if (threadIdx.x < 64) { // thread = (column o2, lane c)
        const int o2 = threadIdx.x >> 2, c = threadIdx.x & 3;
        const uint64_t col2 = (uint64_t)blockIdx.x * 16 + o2;
        if (col2 < m4) {
            unsigned long long t = 0;
#pragma unroll
            for (int gg = 0; gg < G; gg++) t += part[gg][o2][c]; // <= 64 canonical values
            const uint32_t val = bb::reduce64_scaled<0>(t);
            dst[col2 * 4 + c] = val;
            ((volatile unsigned long long *)words)[col2 * 4 + c] = mail_tagged(seq, val);
        }
    }
}
void launch_foldk_wide(const uint32_t *src, uint32_t *dst, uint64_t n, int k, const WideWeights &w, unsigned long long *words,
                       unsigned long long seq, cudaStream_t st) {
    const uint64_t m4 = (n >> k) / 4;
    const int grid = (int)((m4 + 15) / 16);
    switch (k) {
    case 1: k_foldk_wide<1><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    case 2: k_foldk_wide<2><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    case 3: k_foldk_wide<3><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    case 4: k_foldk_wide<4><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    case 5: k_foldk_wide<5><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    case 6: k_foldk_wide<6><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    case 7: k_foldk_wide<7><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    default: k_foldk_wide<8><<<grid, LIN_TPB, 0, st>>>(src, dst, m4, w, words, seq); break;
    }
}

// ---------------------------------------------------------------------------------------------
// Multilinear.eval, LSB-first. One stage folds NV <= 12 variables: a CTA folds a tile of 2^NV consecutive
// elements to one value: 16 elements per thread in registers (4 variables), 5 variables by warp shuffle,
// 3 variables through shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) k_eval_stage(const uint32_t *src, uint64_t n_tiles, int nv, EvalPoint pt,
                                                        uint32_t *out, Mailbox mb, int publish) {
    __shared__ uint32_t sm[THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile_elems = 1ull << nv;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t *base = src + tile * tile_elems;
        uint32_t v = 0;
        int var = 0;
        if (nv >= 4) {
            const uint64_t chunk = (uint64_t)threadIdx.x * 16;
            uint32_t e[16];
            if (chunk < tile_elems) {
                const uint4 *p = reinterpret_cast<const uint4 *>(base + chunk);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint4 q = p[j];
                    e[4 * j] = q.x;
                    e[4 * j + 1] = q.y;
                    e[4 * j + 2] = q.z;
                    e[4 * j + 3] = q.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) e[j] = 0;
            }
#pragma unroll
            for (int lv = 0; lv < 4; lv++) {
#pragma unroll
                for (int j = 0; j < (8 >> lv); j++) e[j] = bb::lerp(e[2 * j], e[2 * j + 1], pt.r[lv], pt.rp[lv]);
            }
            v = e[0];
            var = 4;
        } else {
            // tiny tile (< 16 elements): lane j of warp 0 holds element j, the shuffle tree below does the rest
            v = (threadIdx.x < tile_elems) ? base[threadIdx.x] : 0;
        }
        // warp-level: lane L holds group L; fold adjacent lanes
#pragma unroll
        for (int step = 0; step < 5; step++) {
            uint32_t other = __shfl_down_sync(0xffffffffu, v, 1 << step);
            if (var < nv) {
                v = bb::lerp(v, other, pt.r[var], pt.rp[var]);
                var++;
            }
        }
        if (lane == 0) sm[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t w[THREADS / 32];
#pragma unroll
            for (int j = 0; j < THREADS / 32; j++) w[j] = sm[j];
            int vv = var;
#pragma unroll
            for (int lv = 0; lv < 3; lv++) {
                if (vv < nv) {
#pragma unroll
                    for (int j = 0; j < (4 >> lv); j++) w[j] = bb::lerp(w[2 * j], w[2 * j + 1], pt.r[vv], pt.rp[vv]);
                    vv++;
                }
            }
            out[tile] = w[0];
            if (publish) {
                ((volatile unsigned long long *)mb.mail)[0] = mb.tagged ? mail_tagged(mb.seq, w[0]) : w[0];
                if (!mb.tagged) __threadfence_system();
                ((volatile unsigned long long *)mb.mail)[MAIL_WORDS] = mb.seq;
            }
        }
        __syncthreads();
    }
}

// The last stage(s) in ONE launch: every CTA folds tiles of 2^nv elements (as k_eval_stage does), and the last CTA to finish
// (atomic ticket) folds the n_tiles = 2^nv2 (<= 256) partial results with the remaining variables and publishes the value —
// instead of two launches of ~10 us each for the 2^18 leftovers of a 2^28-entry evaluation.
__device__ __forceinline__ uint32_t cta_fold_tile(const uint32_t *base, uint64_t tile_elems, int nv, const EvalPoint &pt, uint32_t *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v = 0;
    int var = 0;
    if (nv >= 4) {
        const uint64_t chunk = (uint64_t)threadIdx.x * 16;
        uint32_t e[16];
        if (chunk < tile_elems) {
            const uint4 *p = reinterpret_cast<const uint4 *>(base + chunk);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint4 q = __ldcg(p + j);
                e[4 * j] = q.x;
                e[4 * j + 1] = q.y;
                e[4 * j + 2] = q.z;
                e[4 * j + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) e[j] = 0;
        }
#pragma unroll
        for (int lv = 0; lv < 4; lv++) {
#pragma unroll
            for (int j = 0; j < (8 >> lv); j++) e[j] = bb::lerp(e[2 * j], e[2 * j + 1], pt.r[lv], pt.rp[lv]);
        }
        v = e[0];
        var = 4;
    } else {
        v = (threadIdx.x < tile_elems) ? __ldcg(base + threadIdx.x) : 0;
    }
#pragma unroll
    for (int step = 0; step < 5; step++) {
        uint32_t other = __shfl_down_sync(0xffffffffu, v, 1 << step);
        if (var < nv) {
            v = bb::lerp(v, other, pt.r[var], pt.rp[var]);
            var++;
        }
    }
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    uint32_t res = 0;
    if (threadIdx.x == 0) {
        uint32_t w[THREADS / 32];
#pragma unroll
        for (int j = 0; j < THREADS / 32; j++) w[j] = sm[j];
        int vv = var;
#pragma unroll
        for (int lv = 0; lv < 3; lv++) {
            if (vv < nv) {
#pragma unroll
                for (int j = 0; j < (4 >> lv); j++) w[j] = bb::lerp(w[2 * j], w[2 * j + 1], pt.r[vv], pt.rp[vv]);
                vv++;
            }
        }
        res = w[0];
    }
    __syncthreads();
    return res; // valid in thread 0
}

__global__ void __launch_bounds__(THREADS) k_eval_finish(const uint32_t *src, uint64_t n_tiles, int nv, EvalPoint pt, uint32_t *out, int nv2,
                                                         EvalPoint pt2, Mailbox mb) {
    __shared__ uint32_t sm[THREADS / 32];
    __shared__ bool is_last;
    const uint64_t tile_elems = 1ull << nv;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t v = cta_fold_tile(src + tile * tile_elems, tile_elems, nv, pt, sm);
        if (threadIdx.x == 0) out[tile] = v;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(mb.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const uint32_t v = cta_fold_tile(out, n_tiles, nv2, pt2, sm);
    if (threadIdx.x == 0) {
        *mb.ticket = 0u;
        ((volatile unsigned long long *)mb.mail)[0] = mb.tagged ? mail_tagged(mb.seq, v) : v;
        if (!mb.tagged) __threadfence_system();
        ((volatile unsigned long long *)mb.mail)[MAIL_WORDS] = mb.seq;
    }
}

void launch_eval_finish(const uint32_t *src, uint64_t n, int nv, const EvalPoint &pt, uint32_t *out, int nv2, const EvalPoint &pt2,
                        const Mailbox &mb, int sm, cudaStream_t st) {
    const uint64_t n_tiles = n >> nv;
    const uint64_t cap = (uint64_t)sm * 8;
    k_eval_finish<<<(int)(n_tiles < cap ? n_tiles : cap), THREADS, 0, st>>>(src, n_tiles, nv, pt, out, nv2, pt2, mb);
}

// ---- batched evaluation (count polynomials, one point each; blockIdx.y = polynomial) ----
__device__ __forceinline__ void load_point(const uint32_t *pts, uint32_t poly, uint32_t v, uint32_t var0, int nv, EvalPoint &pt) {
    const uint32_t *r = pts + (size_t)poly * 2 * v + var0, *rp = r + v;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        pt.r[k] = k < nv ? r[k] : 0;
        pt.rp[k] = k < nv ? rp[k] : 0;
    }
}
__global__ void __launch_bounds__(THREADS) k_eval_stage_batch(const uint32_t *const *srcs, const uint32_t *src_rows, uint64_t n, int nv,
                                                              const uint32_t *pts, uint32_t v, uint32_t var0, uint32_t *out,
                                                              unsigned long long *publish, Mailbox mb) {
    __shared__ uint32_t sm[THREADS / 32];
    __shared__ bool is_last;
    const uint32_t poly = blockIdx.y;
    EvalPoint pt;
    load_point(pts, poly, v, var0, nv, pt);
    const uint32_t *src = srcs ? srcs[poly] : src_rows + (size_t)poly * n;
    const uint64_t tile_elems = 1ull << nv, n_tiles = n >> nv;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t x = cta_fold_tile(src + tile * tile_elems, tile_elems, nv, pt, sm);
        if (threadIdx.x == 0) {
            out[(size_t)poly * n_tiles + tile] = x;
            if (publish) ((volatile unsigned long long *)publish)[poly] = x;
        }
    }
    if (publish == nullptr) return;
    if (threadIdx.x == 0) {
        __threadfence_system();
        is_last = (atomicAdd(mb.ticket, 1u) == gridDim.x * gridDim.y - 1);
        if (is_last) {
            *mb.ticket = 0u;
            __threadfence_system();
            ((volatile unsigned long long *)mb.mail)[MAIL_WORDS] = mb.seq;
        }
    }
}
void launch_eval_stage_batch(const uint32_t *const *srcs, const uint32_t *src_rows, uint64_t n, uint32_t count, int nv,
                             const uint32_t *pts, uint32_t v, uint32_t var0, uint32_t *out, unsigned long long *publish,
                             const Mailbox &mb, cudaStream_t st) {
    const uint64_t n_tiles = n >> nv;
    dim3 grid((unsigned)(n_tiles < 1024 ? n_tiles : 1024), count);
    k_eval_stage_batch<<<grid, THREADS, 0, st>>>(srcs, src_rows, n, nv, pts, v, var0, out, publish, mb);
}

__global__ void __launch_bounds__(THREADS) k_eval_warp10_batch(const uint32_t *const *srcs, uint64_t n_tiles, const uint32_t *pts, uint32_t v,
                                                               uint32_t var0, uint32_t *out) {
    const uint32_t poly = blockIdx.y;
    EvalPoint pt;
    load_point(pts, poly, v, var0, 10, pt);
    const uint32_t *__restrict__ src = srcs[poly];
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * THREADS) >> 5;
    for (uint64_t t = warp; t < n_tiles; t += n_warps) {
        const uint4 *p = reinterpret_cast<const uint4 *>(src + t * 1024) + lane;
        uint4 q[8];
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = __ldg(p + 32 * j);
        uint32_t e[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t a = bb::lerp(q[j].x, q[j].y, pt.r[0], pt.rp[0]);
            const uint32_t b = bb::lerp(q[j].z, q[j].w, pt.r[0], pt.rp[0]);
            e[j] = bb::lerp(a, b, pt.r[1], pt.rp[1]);
        }
#pragma unroll
        for (int lv = 0; lv < 3; lv++) {
#pragma unroll
            for (int j = 0; j < (4 >> lv); j++) e[j] = bb::lerp(e[2 * j], e[2 * j + 1], pt.r[7 + lv], pt.rp[7 + lv]);
        }
        uint32_t x = e[0];
#pragma unroll
        for (int step = 0; step < 5; step++) {
            uint32_t other = __shfl_down_sync(0xffffffffu, x, 1 << step);
            x = bb::lerp(x, other, pt.r[2 + step], pt.rp[2 + step]);
        }
        if (lane == 0) out[(size_t)poly * n_tiles + t] = x;
    }
}
void launch_eval_warp10_batch(const uint32_t *const *srcs, uint64_t n, uint32_t count, const uint32_t *pts, uint32_t v, uint32_t var0,
                              uint32_t *out, int sm, cudaStream_t st) {
    const uint64_t n_tiles = n >> 10;
    uint64_t ctas = (n_tiles + THREADS / 32 - 1) / (THREADS / 32), cap = ((uint64_t)sm * 4 + count - 1) / count;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    dim3 grid((unsigned)ctas, count);
    k_eval_warp10_batch<<<grid, THREADS, 0, st>>>(srcs, n_tiles, pts, v, var0, out);
}

// Big stages: one WARP folds a tile of 1024 consecutive elements (10 variables) with fully coalesced 512-byte
// loads and no block barrier. The multilinear extension is symmetric in the order variables are bound (exact
// arithmetic), so the kernel binds them in the order the data arrives: bits 0,1 (inside a uint4), bits 7,8,9 (the
// thread's 8 loads, 128 elements apart), then bits 2..6 (lanes, by shuffle). UT tiles are in flight per warp.
template <int UT>
__global__ void __launch_bounds__(THREADS) k_eval_warp10(const uint32_t *__restrict__ src, uint64_t n_tiles, EvalPoint pt,
                                                         uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * THREADS) >> 5;
    for (uint64_t t0 = warp * UT; t0 < n_tiles; t0 += n_warps * UT) {
        uint4 v[UT][8];
#pragma unroll
        for (int u = 0; u < UT; u++) {
            if (t0 + u < n_tiles) {
                const uint4 *p = reinterpret_cast<const uint4 *>(src + (t0 + u) * 1024) + lane;
#pragma unroll
                for (int j = 0; j < 8; j++) v[u][j] = __ldg(p + 32 * j);
            }
        }
#pragma unroll
        for (int u = 0; u < UT; u++) {
            if (t0 + u < n_tiles) {
                uint32_t e[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t a = bb::lerp(v[u][j].x, v[u][j].y, pt.r[0], pt.rp[0]);
                    uint32_t b = bb::lerp(v[u][j].z, v[u][j].w, pt.r[0], pt.rp[0]);
                    e[j] = bb::lerp(a, b, pt.r[1], pt.rp[1]);
                }
#pragma unroll
                for (int lv = 0; lv < 3; lv++) {
#pragma unroll
                    for (int j = 0; j < (4 >> lv); j++) e[j] = bb::lerp(e[2 * j], e[2 * j + 1], pt.r[7 + lv], pt.rp[7 + lv]);
                }
                uint32_t x = e[0];
#pragma unroll
                for (int step = 0; step < 5; step++) {
                    uint32_t other = __shfl_down_sync(0xffffffffu, x, 1 << step);
                    x = bb::lerp(x, other, pt.r[2 + step], pt.rp[2 + step]);
                }
                if (lane == 0) out[t0 + u] = x;
            }
        }
    }
}

// The same stage with the tile fetched by ONE bulk copy per warp (cp.async.bulk, 4 KB) into a per-warp ring of shared-memory
// stages; every warp is its own producer (lane 0 issues the copy for the tile ES stages ahead) and consumer (all lanes wait
// on the stage's mbarrier, then read their 8 uint4 with conflict-free LDS.128): no register-staged loads, no block barrier,
// ES x 4 KB per warp in flight at all times.
template <int EVB_WARPS, int EVB_STAGES>
__global__ void __launch_bounds__(EVB_WARPS * 32) k_eval_warp10_bulk(const uint32_t *__restrict__ src, uint64_t n_tiles, EvalPoint pt,
                                                                        uint32_t *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint4 *ring = reinterpret_cast<uint4 *>(smem_raw) + (size_t)wib * EVB_STAGES * 256; // [stage][256 uint4]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)EVB_WARPS * EVB_STAGES * 4096) + wib * EVB_STAGES;
    if (lane == 0) {
        for (int st = 0; st < EVB_STAGES; st++) blk::mbar_init(&full[st], 1);
        blk::fence_barrier_init();
    }
    __syncwarp();
    const uint64_t warp = (uint64_t)blockIdx.x * EVB_WARPS + wib, n_warps = (uint64_t)gridDim.x * EVB_WARPS;
    const uint64_t cnt = warp < n_tiles ? (n_tiles - warp + n_warps - 1) / n_warps : 0;
    auto issue = [&](uint64_t it) { // lane 0 only
        const int st = (int)(it % EVB_STAGES);
        blk::mbar_expect_tx(&full[st], 4096);
        blk::g2s(ring + (size_t)st * 256, src + (warp + it * n_warps) * 1024, 4096, &full[st]);
    };
    if (lane == 0)
        for (uint64_t it = 0; it < (uint64_t)(EVB_STAGES - 1) && it < cnt; it++) issue(it);
    for (uint64_t it = 0; it < cnt; it++) {
        const int st = (int)(it % EVB_STAGES);
        if (lane == 0 && it + EVB_STAGES - 1 < cnt) {
            blk::fence_proxy_async(); // the stage being refilled was read (generic proxy) one iteration ago
            issue(it + EVB_STAGES - 1);
        }
        blk::mbar_wait(&full[st], (uint32_t)(it / EVB_STAGES) & 1);
        const uint4 *tile = ring + (size_t)st * 256 + lane;
        uint32_t e[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint4 v = tile[32 * j];
            const uint32_t a = bb::lerp(v.x, v.y, pt.r[0], pt.rp[0]);
            const uint32_t b = bb::lerp(v.z, v.w, pt.r[0], pt.rp[0]);
            e[j] = bb::lerp(a, b, pt.r[1], pt.rp[1]);
        }
#pragma unroll
        for (int lv = 0; lv < 3; lv++) {
#pragma unroll
            for (int j = 0; j < (4 >> lv); j++) e[j] = bb::lerp(e[2 * j], e[2 * j + 1], pt.r[7 + lv], pt.rp[7 + lv]);
        }
        uint32_t x = e[0];
#pragma unroll
        for (int step = 0; step < 5; step++) {
            uint32_t other = __shfl_down_sync(0xffffffffu, x, 1 << step);
            x = bb::lerp(x, other, pt.r[2 + step], pt.rp[2 + step]);
        }
        if (lane == 0) out[warp + it * n_warps] = x;
        // the shuffles above are warp-wide: every lane has consumed its LDS results before lane 0 refills this stage
    }
}

void launch_eval_warp10(const uint32_t *src, uint64_t n, const EvalPoint &pt, uint32_t *out, int sm, cudaStream_t st) {
    const uint64_t n_tiles = n >> 10;
    // measured at 2^28 (profiles/r02_sweep2.txt): plain loads 6250 GB/s; bulk ring 8 warps x 6 stages 4698, 16 x 3 6654,
    // 4 x 6 (2 CTAs/SM) 4753, 8 x 3 (2 CTAs/SM) 6444, 32 x 1 6705, 16 x 2 6848 <- default
    static const int BULK = tune("ZB_EVAL_BULK", 6);
    auto bulk = [&](auto warps_c, auto stages_c, int per_sm) {
        constexpr int W = decltype(warps_c)::value, S = decltype(stages_c)::value;
        constexpr int SMEM = W * S * 4096 + W * S * 8;
        ENSURE_DYN_SMEM((k_eval_warp10_bulk<W, S>), SMEM);
        uint64_t ctas = (n_tiles + W - 1) / W;
        if (ctas > (uint64_t)sm * per_sm) ctas = (uint64_t)sm * per_sm;
        k_eval_warp10_bulk<W, S><<<(int)(ctas ? ctas : 1), W * 32, SMEM, st>>>(src, n_tiles, pt, out);
    };
    using std::integral_constant;
    if (BULK == 1) return bulk(integral_constant<int, 8>{}, integral_constant<int, 6>{}, 1);
    if (BULK == 2) return bulk(integral_constant<int, 16>{}, integral_constant<int, 3>{}, 1);
    if (BULK == 3) return bulk(integral_constant<int, 4>{}, integral_constant<int, 6>{}, 2);
    if (BULK == 4) return bulk(integral_constant<int, 8>{}, integral_constant<int, 3>{}, 2);
    if (BULK == 5) return bulk(integral_constant<int, 32>{}, integral_constant<int, 1>{}, 1);
    if (BULK == 6) return bulk(integral_constant<int, 16>{}, integral_constant<int, 2>{}, 1);
    static const int UT = tune("ZB_EVAL_UT", 2);
    static const int CPS = tune("ZB_EVAL_CPS", 2);
    const uint64_t warps_needed = (n_tiles + UT - 1) / UT;
    uint64_t ctas = (warps_needed + THREADS / 32 - 1) / (THREADS / 32);
    const uint64_t cap = (uint64_t)sm * CPS;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    if (UT >= 4) k_eval_warp10<4><<<(int)ctas, THREADS, 0, st>>>(src, n_tiles, pt, out);
    else if (UT >= 2) k_eval_warp10<2><<<(int)ctas, THREADS, 0, st>>>(src, n_tiles, pt, out);
    else k_eval_warp10<1><<<(int)ctas, THREADS, 0, st>>>(src, n_tiles, pt, out);
}

void launch_eval_stage(const uint32_t *src, uint64_t n, int nvars, const EvalPoint &pt, uint32_t *out, const Mailbox *mb,
                       int sm, cudaStream_t st) {
    uint64_t n_tiles = n >> nvars;
    uint64_t cap = (uint64_t)sm * 8;
    int grid = (int)(n_tiles < cap ? n_tiles : cap);
    Mailbox m = mb ? *mb : Mailbox{nullptr, nullptr, nullptr, 0, nullptr, 0, false};
    k_eval_stage<<<grid, THREADS, 0, st>>>(src, n_tiles, nvars, pt, out, m, mb != nullptr);
}

// ---------------------------------------------------------------------------------------------
// element-wise helpers
// ---------------------------------------------------------------------------------------------
__global__ void k_narrow(const uint64_t *src, uint32_t *dst, uint64_t n, unsigned int *err) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t v = src[i];
        bad |= (v >= bb::P);
        dst[i] = (uint32_t)v;
    }
    if (bad) atomicOr(err, 1u);
}
__global__ void k_check(const uint32_t *src, uint64_t n, unsigned int *err) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= (src[i] >= bb::P);
    if (bad) atomicOr(err, 1u);
}
__global__ void k_widen(const uint32_t *src, uint64_t *dst, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
__global__ void k_fill(uint32_t *dst, uint64_t n, uint32_t v) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void k_synthetic(uint32_t *dst, uint64_t n, uint64_t seed, uint64_t start, uint64_t step) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = bb::reduce64_scaled<0>(splitmix64(seed + start + i * step));
}
// eq table of a point from two half tables: out[i] = lo[i & (2^s - 1)] * hi[i >> s], hi in Montgomery form (one product per entry;
// the halves have <= 2^16 entries and stay in L1/L2)
__global__ void k_eq_table(const uint32_t *__restrict__ lo, const uint32_t *__restrict__ hi_mont, int s, uint64_t n, uint32_t *out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, mask = (1ull << s) - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = bb::mont_mul(__ldg(lo + (i & mask)), __ldg(hi_mont + (i >> s)));
}
__global__ void k_add(const uint32_t *a, const uint32_t *b, uint32_t *o, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) o[i] = bb::add(a[i], b[i]);
}
__global__ void k_scalar_mul(const uint32_t *a, uint32_t s, uint32_t sp, uint32_t *o, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) o[i] = bb::mul_shoup(a[i], s, sp);
}

// WitnessGenerator.generate on SoA trace columns — /root/reference/src/constraints/witness.zig:29-270.
// One launch packs a chunk [step0, step0 + chunk) of all columns: value mod p (F.init), and, for the part of the
// chunk past num_steps, the padding rule of the column (first n_hold columns repeat the last real value, the others
// are zero). cols: chunk-local staging, column-major [n_cols][chunk]; last_vals: per column F.init(last real value).
struct WitnessOut {
    uint32_t *col[64];
};
__global__ void k_witness_pack(const uint64_t *cols, uint64_t chunk, uint64_t step0, uint64_t num_steps, uint64_t padded,
                               uint32_t n_hold, const uint32_t *last_vals, WitnessOut out) {
    const uint32_t c = blockIdx.y;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t *o = out.col[c];
    const uint32_t fill = c < n_hold ? last_vals[c] : 0u;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < chunk; j += stride) {
        const uint64_t i = step0 + j;
        if (i >= padded) break;
        o[i] = i < num_steps ? bb::reduce64_scaled<0>(cols[c * chunk + j]) : fill; // F.init: any u64 mod p
    }
}
void launch_witness_pack(const uint64_t *cols, uint64_t chunk, uint64_t step0, uint64_t num_steps, uint64_t padded,
                         uint32_t n_cols, uint32_t n_hold, const uint32_t *last_vals, uint32_t *const *out_cols, cudaStream_t st) {
    WitnessOut w{};
    for (uint32_t c = 0; c < n_cols; c++) w.col[c] = out_cols[c];
    uint64_t g = (chunk + 255) / 256;
    dim3 grid((unsigned)(g > 148 * 4 ? 148 * 4 : (g ? g : 1)), n_cols);
    k_witness_pack<<<grid, 256, 0, st>>>(cols, chunk, step0, num_steps, padded, n_hold, last_vals, w);
}

static inline int ew_grid(uint64_t n) {
    uint64_t g = (n + 255) / 256;
    if (g < 1) g = 1;
    return (int)(g > 148 * 16 ? 148 * 16 : g);
}
void launch_narrow_u64(const uint64_t *src, uint32_t *dst, uint64_t n, unsigned int *err, cudaStream_t st) {
    k_narrow<<<ew_grid(n), 256, 0, st>>>(src, dst, n, err);
}
void launch_check_u32(const uint32_t *src, uint64_t n, unsigned int *err, cudaStream_t st) {
    k_check<<<ew_grid(n), 256, 0, st>>>(src, n, err);
}
void launch_widen_u32(const uint32_t *src, uint64_t *dst, uint64_t n, cudaStream_t st) {
    k_widen<<<ew_grid(n), 256, 0, st>>>(src, dst, n);
}
void launch_fill(uint32_t *dst, uint64_t n, uint32_t v, cudaStream_t st) { k_fill<<<ew_grid(n), 256, 0, st>>>(dst, n, v); }
void launch_synthetic(uint32_t *dst, uint64_t n, uint64_t seed, uint64_t start, uint64_t step, cudaStream_t st) {
    k_synthetic<<<ew_grid(n), 256, 0, st>>>(dst, n, seed, start, step);
}
void launch_eq_table(const uint32_t *lo, const uint32_t *hi_mont, int s, uint64_t n, uint32_t *out, cudaStream_t st) {
    k_eq_table<<<ew_grid(n), 256, 0, st>>>(lo, hi_mont, s, n, out);
}
void launch_add(const uint32_t *a, const uint32_t *b, uint32_t *o, uint64_t n, cudaStream_t st) {
    k_add<<<ew_grid(n), 256, 0, st>>>(a, b, o, n);
}
void launch_scalar_mul(const uint32_t *a, uint32_t s, uint32_t *o, uint64_t n, cudaStream_t st) {
    k_scalar_mul<<<ew_grid(n), 256, 0, st>>>(a, s, bb::shoup_pre(s), o, n);
}

// ---------------------------------------------------------------------------------------------
// Lasso row hashing: hashEntry / hashQuery — /root/reference/src/lookups/lasso_prover.zig:208-239
// XXH3-64 of an 8-byte input with seed 0 (the 4..8-byte path of the XXH3 spec)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
__device__ __forceinline__ uint64_t xxh3_8(uint64_t h) {
    // input64 = in2 + (in1 << 32) with in1 = low word, in2 = high word of le64(h): i.e. the words swapped
    const uint64_t bitflip = 0x1cad21f72c81017cull ^ 0xdb979083e96dd4deull; // secret[8..16) ^ secret[16..24), seed 0
    uint64_t x = ((h >> 32) | (h << 32)) ^ bitflip;
    x ^= rotl64(x, 49) ^ rotl64(x, 24);
    x *= 0x9FB21C651E98DF25ull;
    x ^= (x >> 35) + 8;
    x *= 0x9FB21C651E98DF25ull;
    return x ^ (x >> 28);
}
__device__ __forceinline__ uint32_t hash_row3(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t h = xxh3_8(a);
    h = xxh3_8(h ^ b);
    h = xxh3_8(h ^ c);
    return bb::reduce64_scaled<0>(h);
}

__global__ void k_xxh3_rows(const uint32_t *rows, uint64_t n_rows, uint32_t arity, uint64_t n_padded, uint32_t *out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_padded; i += stride) {
        uint32_t v = 0;
        if (i < n_rows) {
            uint64_t h = 0;
            for (uint32_t k = 0; k < arity; k++) h = xxh3_8(h ^ (uint64_t)rows[i * arity + k]);
            v = bb::reduce64_scaled<0>(h);
        }
        out[i] = v;
    }
}
void launch_xxh3_rows(const uint32_t *rows, uint64_t n_rows, uint32_t arity, uint64_t n_padded, uint32_t *out, cudaStream_t st) {
    k_xxh3_rows<<<ew_grid(n_padded), 256, 0, st>>>(rows, n_rows, arity, n_padded, out);
}

// buildAddTable / buildXorTable / buildAndTable (table_builder.zig:126-213) generated and hashed in registers
__global__ void k_table_mle(int op, uint32_t bits, uint32_t *out) {
    const uint64_t n = 1ull << (2 * bits);
    const uint64_t mask = (1ull << bits) - 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t a = i >> bits, b = i & mask;
        uint64_t r = op == 0 ? ((a + b) & mask) : op == 1 ? (a ^ b) : (a & b);
        out[i] = hash_row3(a % bb::P, b % bb::P, r % bb::P);
    }
}
void launch_table_mle(int op, uint32_t bits, uint32_t *out, cudaStream_t st) {
    k_table_mle<<<ew_grid(1ull << (2 * bits)), 256, 0, st>>>(op, bits, out);
}

} // namespace zk
