// Host twin of the zigz prover-side API (include/zigz_host.h). Everything here runs on the CPU and reaches the
// GPU only through the public device C ABI (zb_* of include/zigz_b200.h): this file is what the reference's Zig
// bodies become once their hot loops are replaced by extern calls (INTEGRATION.md shows the Zig side).
#include "../../include/zigz_host.h"
#include "sha3_host.hpp"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <memory>
#include <new>
#include <thread>
#include <vector>

using zigz::Sha3_256;

namespace {
constexpr uint64_t P = ZB_BABYBEAR_P;

inline uint64_t f_add(uint64_t a, uint64_t b) { // field.zig:73-88
    uint64_t s = a + b;
    return s >= P ? s - P : s;
}
inline uint64_t f_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : P - (b - a); } // field.zig:91-98
// field.zig:112-147 computes (u128)a*b % p; for canonical operands (< p < 2^31) the product fits 62 bits, so the same
// canonical value comes out of a plain 64-bit remainder (a multiply-by-reciprocal, ~10x faster than the __umodti3 call)
inline uint64_t f_mul(uint64_t a, uint64_t b) { return (a * b) % P; }

inline uint64_t digest_to_field(const uint8_t d[32]) { // hash.zig:228-242: first 8 bytes LE, then F.init = mod p
    uint64_t v;
    memcpy(&v, d, 8);
    return v % P;
}
} // namespace

struct zh_transcript {
    Sha3_256 h;
};

extern "C" {

zh_transcript *zh_transcript_new(void) { return new (std::nothrow) zh_transcript(); }
zh_transcript *zh_transcript_clone(const zh_transcript *t) { return t ? new (std::nothrow) zh_transcript(*t) : nullptr; }
void zh_transcript_free(zh_transcript *t) { delete t; }
void zh_transcript_append_field(zh_transcript *t, uint64_t v) { t->h.update_words(&v, 1); }
void zh_transcript_append_fields(zh_transcript *t, const uint64_t *v, size_t n) { t->h.update_words(v, n); }
void zh_transcript_append_bytes(zh_transcript *t, const void *d, size_t n) { t->h.update(d, n); }
uint64_t zh_transcript_challenge(zh_transcript *t) { // hash.zig:301-316
    uint8_t digest[32];
    t->h.peek(digest);
    uint64_t r = digest_to_field(digest);
    t->h.update(digest, 32); // the transcript absorbs its own digest
    return r;
}
void zh_transcript_finalize(zh_transcript *t, uint8_t out[32]) { t->h.peek(out); }
void zh_sha3_256(const void *d, size_t n, uint8_t out[32]) { Sha3_256::hash(d, n, out); }
uint64_t zh_digest_to_field(const uint8_t digest[32]) { return digest_to_field(digest); }

uint64_t zh_f_add(uint64_t a, uint64_t b) { return f_add(a % P, b % P); }
uint64_t zh_f_sub(uint64_t a, uint64_t b) { return f_sub(a % P, b % P); }
uint64_t zh_f_mul(uint64_t a, uint64_t b) { return f_mul(a % P, b % P); }

uint64_t zh_eval_univariate(const uint64_t *c, uint32_t n, uint64_t x) { // sumcheck_protocol.zig:113-123 (Horner)
    if (n == 0) return 0;
    x %= P; // F.init
    uint64_t r = c[n - 1] % P;
    for (uint32_t i = n - 1; i > 0; i--) r = f_add(f_mul(r, x), c[i - 1] % P);
    return r;
}

/* ------------------------------------------------------------------ sumcheck */

// The round loop of SumcheckProver.prove for d polynomials (d == 1: sumcheck_prover.zig:50-77).
//   device: round coefficients  ->  host: absorb + challenge  ->  device: fold (fused with the next round's sums)
// `consume`: fold the caller's polynomials in place; otherwise the first fold goes to fresh buffers (the
// reference's copy at :47 costs a full pass; folding out of place in round 0 gives the same isolation for free).
//
// Multi-GPU (a communicator is attached to ctx, world = P): `polys` are this rank's CYCLIC shards (local element j
// is global index rank + P*j), so MSB-first pairs never cross GPUs while the shards have >= 2 entries.
//   phase 1 (big tables): one launch per round; the kernel's partial sums are all-reduced over NCCL in stream order
//            ("comm_reduce") and every rank's host runs the same transcript on the same coefficients;
//   phase 2 (shards <= 2^ZB_GATHER_LOG2 entries, default 2^16): the shards are all-gathered once into the global order
//            and every GPU finishes the remaining rounds on the whole (small) table, redundantly and without any
//            further exchange — latency-bound rounds should not pay a collective each.
// evaluations on the point set of degree d ({0,1} / {0,1,inf} / {0,1,-1,inf}) -> coefficients [a0..ad]
static void evals_to_coeffs(uint32_t d, const uint64_t *e, uint64_t *c) {
    const uint64_t half = (P + 1) / 2;
    if (d == 1) {
        c[0] = e[0];
        c[1] = f_sub(e[1], e[0]); // multilinear.zig:229
    } else if (d == 2) {
        c[0] = e[0];
        c[2] = e[2];
        c[1] = f_sub(f_sub(e[1], e[0]), e[2]);
    } else {
        c[0] = e[0];
        c[3] = e[3];
        c[2] = f_sub(f_mul(f_add(e[1], e[2]), half), e[0]);
        c[1] = f_sub(f_mul(f_sub(e[1], e[2]), half), e[3]);
    }
}

// Round polynomials out of the bivariate grid G(X, Y) of zb_prod_grid / zb_prod_fold_grid:
//   this round:  g(X)  = G(X, 0) + G(X, 1)
//   next round:  g'(Y) = G(r, Y), r = this round's challenge (interpolate every column in X, evaluate at r)
static void grid_round_a(uint32_t d, const uint64_t *grid, uint64_t *coeffs) {
    const uint32_t np = d == 1 ? 2 : d + 1;
    uint64_t e[4] = {0, 0, 0, 0};
    for (uint32_t ix = 0; ix < np; ix++) e[ix] = f_add(grid[ix * np + 0], grid[ix * np + 1]);
    evals_to_coeffs(d, e, coeffs);
}
static void grid_round_b(uint32_t d, const uint64_t *grid, uint64_t r, uint64_t *coeffs) {
    const uint32_t np = d == 1 ? 2 : d + 1;
    uint64_t ey[4] = {0, 0, 0, 0};
    for (uint32_t iy = 0; iy < np; iy++) {
        uint64_t col[4] = {0, 0, 0, 0}, cx[4] = {0, 0, 0, 0};
        for (uint32_t ix = 0; ix < np; ix++) col[ix] = grid[ix * np + iy];
        evals_to_coeffs(d, col, cx);
        // (the Y = inf column holds the leading Y-coefficient of G: a degree-d polynomial in X like the other columns)
        ey[iy] = zh_eval_univariate(cx, d + 1, r);
    }
    evals_to_coeffs(d, ey, coeffs);
}

// process-wide tuning knob (test hook zh_set_grid_min_log2): atomic, so that host threads driving different contexts
// (one per GPU of a device-mask context) can read it while a test thread sets it
static std::atomic<int> g_grid_min_log2{-1}; // -1: the default below
// folded tables below 2^this are bound one round per kernel (and by the persistent tail); 0 disables the grid path. Default: 5
// when small tables finish on the host anyway ("prod_host_tail_log2" > 0: every device pass then serves two rounds), else 15
// (below that the persistent single-round tail is the cheaper way down)
static int grid_min_log2(bool host_tail_on) {
    const int v = g_grid_min_log2.load(std::memory_order_relaxed);
    if (v >= 0) return v;
    static const int env = [] {
        const char *e = getenv("ZB_GRID_MIN_LOG2");
        return e && *e ? (atoi(e) < 0 ? 0 : atoi(e)) : -1;
    }();
    return env >= 0 ? env : (host_tail_on ? 5 : 15);
}

static int gather_log2() {
    static const int v = [] {
        const char *e = getenv("ZB_GATHER_LOG2");
        int x = e && *e ? atoi(e) : 16;
        return x < 1 ? 1 : (x > 24 ? 24 : x); // >= 2 entries: the round coefficients in hand must still be sums
    }();
    return v;
}

// d = 1 on one GPU, tables of more than 2^10 entries: several rounds per pass over the data (zb_mle_block_sums /
// zb_mle_fold_multi). The 2^k block sums S in hand ARE a 2^k-entry multilinear table whose sumcheck rounds equal the next k
// rounds of the big table (roundPolynomial is linear, multilinear.zig:205-232), so the host runs those rounds on S exactly as
// sumcheck_prover.zig:50-77 does — roundPolynomial, transcript, partialEval — while the device only sees one launch per k
// rounds. When the folded table is small enough to be published whole, S is the table itself and the host finishes the proof.
// ---- the last rounds of a product sumcheck on the host ----
// Four field elements per AVX2 register (one per 64-bit lane, values < 2^32). Montgomery product a b 2^-32 mod p in the
// subtractive form the kernels use (bb.cuh: mont_mul_lazy): with m = lo(t) p^-1 mod 2^32 the difference t - m p is an exact
// multiple of 2^32, so the result is hi(t) - hi(m p) in (-p, p), made canonical with one conditional addition. Needs a b < 2^32 p.
namespace hv {
const uint64_t R1 = (1ull << 32) % P;  // 2^32 mod p
const uint64_t R2 = (R1 * R1) % P;     // 2^64 mod p
const uint64_t R3 = (R2 * R1) % P;     // 2^96 mod p
inline __m256i bcast(uint64_t x) { return _mm256_set1_epi64x((long long)x); }
inline __m256i load4(const uint32_t *p) { return _mm256_cvtepu32_epi64(_mm_loadu_si128((const __m128i *)p)); }
inline void store4(uint32_t *p, __m256i v) { // low 32 bits of every lane
    const __m256i q = _mm256_permutevar8x32_epi32(v, _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0));
    _mm_storeu_si128((__m128i *)p, _mm256_castsi256_si128(q));
}
inline __m256i mont(__m256i a, __m256i b) {
    const __m256i t = _mm256_mul_epu32(a, b);
    const __m256i m = _mm256_mul_epu32(t, bcast(0x88000001u)); // low halves: lo(t) * p^-1 (only the low 32 bits are used next)
    const __m256i mp = _mm256_mul_epu32(m, bcast(P));
    const __m256i u = _mm256_sub_epi64(_mm256_srli_epi64(t, 32), _mm256_srli_epi64(mp, 32));
    return _mm256_add_epi64(u, _mm256_and_si256(_mm256_cmpgt_epi64(_mm256_setzero_si256(), u), bcast(P)));
}
inline __m256i sub(__m256i a, __m256i b) { // canonical a - b mod p
    const __m256i u = _mm256_sub_epi64(a, b);
    return _mm256_add_epi64(u, _mm256_and_si256(_mm256_cmpgt_epi64(_mm256_setzero_si256(), u), bcast(P)));
}
inline __m256i add(__m256i a, __m256i b) { // canonical a + b mod p
    const __m256i s = _mm256_add_epi64(a, b);
    return _mm256_sub_epi64(s, _mm256_andnot_si256(_mm256_cmpgt_epi64(bcast(P), s), bcast(P)));
}
inline uint64_t hsum(__m256i v) { // lanes hold sums of < 2^12 canonical terms: no overflow
    alignas(32) uint64_t x[4];
    _mm256_store_si256((__m256i *)x, v);
    return x[0] + x[1] + x[2] + x[3];
}
} // namespace hv

// T[k][0 .. m) are the d current tables (canonical u32, as published by zb_prod_fold_dump). Each round is roundPolynomial over
// MSB-first pairs — the same point sets as the kernels: {s0, s1} (d = 1), {g(0), g(1), g(inf)} (d = 2), {g(0), g(1), g(-1),
// g(inf)} (d = 3), exact field arithmetic, so the coefficients equal the device path's bit for bit — then the transcript, then
// partialEval (sumcheck_prover.zig:50-77, multilinear.zig:166-173). Halves of >= 4 pairs run four lanes wide.
static void finish_rounds_on_host(uint32_t d, uint32_t *T, uint64_t m, uint32_t round, zh_transcript *tr, const uint64_t *fixed_challenges,
                                  uint64_t *round_polys, uint64_t *final_point, uint64_t *final_evals, uint64_t *claimed_sum) {
    const uint32_t nc = d + 1;
    uint32_t *t0 = T, *t1 = T + m, *t2 = T + 2 * m;
    for (uint64_t len = m; len >= 2; len /= 2, round++) {
        const uint64_t h = len / 2;
        uint64_t e[4] = {0, 0, 0, 0}; // u64 sums of canonical terms (< 2^31 each, <= 2^11 of them): exact
        if (d == 1) {
            for (uint64_t i = 0; i < h; i++) {
                e[0] += t0[i];
                e[1] += t0[i + h];
            }
        } else if (h >= 4) {
            // products come out scaled by 2^-32 per Montgomery step: the sums are rescaled once (R^(d-1)) after the loop
            __m256i s0 = _mm256_setzero_si256(), s1 = s0, s2 = s0, s3 = s0;
            if (d == 2) {
                for (uint64_t i = 0; i < h; i += 4) {
                    const __m256i a0 = hv::load4(t0 + i), b0 = hv::load4(t0 + i + h), a1 = hv::load4(t1 + i), b1 = hv::load4(t1 + i + h);
                    s0 = _mm256_add_epi64(s0, hv::mont(a0, a1));
                    s1 = _mm256_add_epi64(s1, hv::mont(b0, b1));
                    s2 = _mm256_add_epi64(s2, hv::mont(hv::sub(b0, a0), hv::sub(b1, a1)));
                }
                e[0] = f_mul(hv::hsum(s0) % P, hv::R1);
                e[1] = f_mul(hv::hsum(s1) % P, hv::R1);
                e[2] = f_mul(hv::hsum(s2) % P, hv::R1);
            } else {
                for (uint64_t i = 0; i < h; i += 4) {
                    const __m256i a0 = hv::load4(t0 + i), b0 = hv::load4(t0 + i + h), a1 = hv::load4(t1 + i), b1 = hv::load4(t1 + i + h),
                                  a2 = hv::load4(t2 + i), b2 = hv::load4(t2 + i + h);
                    const __m256i d0 = hv::sub(b0, a0), d1 = hv::sub(b1, a1), d2 = hv::sub(b2, a2); // slope = value at infinity
                    s0 = _mm256_add_epi64(s0, hv::mont(hv::mont(a0, a1), a2));
                    s1 = _mm256_add_epi64(s1, hv::mont(hv::mont(b0, b1), b2));
                    s2 = _mm256_add_epi64(s2, hv::mont(hv::mont(hv::sub(a0, d0), hv::sub(a1, d1)), hv::sub(a2, d2))); // X = -1
                    s3 = _mm256_add_epi64(s3, hv::mont(hv::mont(d0, d1), d2));
                }
                e[0] = f_mul(hv::hsum(s0) % P, hv::R2);
                e[1] = f_mul(hv::hsum(s1) % P, hv::R2);
                e[2] = f_mul(hv::hsum(s2) % P, hv::R2);
                e[3] = f_mul(hv::hsum(s3) % P, hv::R2);
            }
        } else if (d == 2) {
            for (uint64_t i = 0; i < h; i++) {
                const uint64_t a0 = t0[i], b0 = t0[i + h], a1 = t1[i], b1 = t1[i + h];
                e[0] += f_mul(a0, a1);
                e[1] += f_mul(b0, b1);
                e[2] += f_mul(f_sub(b0, a0), f_sub(b1, a1));
            }
        } else {
            for (uint64_t i = 0; i < h; i++) {
                const uint64_t a0 = t0[i], b0 = t0[i + h], a1 = t1[i], b1 = t1[i + h], a2 = t2[i], b2 = t2[i + h];
                const uint64_t d0 = f_sub(b0, a0), d1 = f_sub(b1, a1), d2 = f_sub(b2, a2);
                e[0] += f_mul(f_mul(a0, a1), a2);
                e[1] += f_mul(f_mul(b0, b1), b2);
                e[2] += f_mul(f_mul(f_sub(a0, d0), f_sub(a1, d1)), f_sub(a2, d2)); // X = -1: a - (b - a)
                e[3] += f_mul(f_mul(d0, d1), d2);
            }
        }
        for (uint32_t k = 0; k < 4; k++) e[k] %= P;
        uint64_t c[4];
        evals_to_coeffs(d, e, c);
        for (uint32_t k = 0; k < nc; k++) round_polys[(size_t)round * nc + k] = c[k];
        if (round == 0 && claimed_sum) *claimed_sum = f_add(e[0], e[1]); // g(0) + g(1) (sumOverHypercube for d == 1, :40)
        uint64_t r;
        if (fixed_challenges) {
            r = fixed_challenges[round]; // proveInteractive :127
        } else {
            zh_transcript_append_fields(tr, c, nc); // generateChallenge, sumcheck_protocol.zig:176-184
            r = zh_transcript_challenge(tr);
        }
        final_point[round] = r;
        const __m256i rR = hv::bcast(f_mul(r, hv::R1)); // r in Montgomery form: mont(rR, x) = r x
        for (uint32_t k = 0; k < d; k++) { // partialEval :166-173
            uint32_t *t = T + (size_t)k * m;
            if (h >= 4) {
                for (uint64_t i = 0; i < h; i += 4) {
                    const __m256i lo = hv::load4(t + i), hi = hv::load4(t + i + h);
                    hv::store4(t + i, hv::add(lo, hv::mont(hv::sub(hi, lo), rR)));
                }
            } else {
                for (uint64_t i = 0; i < h; i++) t[i] = (uint32_t)f_add(t[i], f_mul(r, f_sub(t[i + h], t[i])));
            }
        }
    }
    for (uint32_t k = 0; k < d; k++) final_evals[k] = T[(size_t)k * m]; // current_poly.evaluations[0] (:88)
}

static int32_t prove_linear(zb_ctx *ctx, zb_mle poly, uint32_t v, bool consume, const uint64_t *fixed_challenges,
                            uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint64_t *claimed_sum) {
    int64_t dump_log2 = 12, kk = 5;
    zb_get_option(ctx, "host_tail_log2", &dump_log2);
    zb_get_option(ctx, "linear_k", &kk);
    const uint32_t K = (uint32_t)kk;    // variables per device pass (large tables)
    alignas(32) uint64_t S[1u << 12];
    alignas(32) uint32_t T[1u << 12];
    uint64_t rs[10];
    uint32_t u = v;                     // log2 of the current device table
    // Small tables (at most 8 variables above the published size): ONE block-sum pass with 2^(v - dump) blocks, then ONE fold
    // that binds all of them and publishes the table — two device round trips for a 2^20-entry prove. Otherwise K per pass.
    uint32_t have = (v > (uint32_t)dump_log2 && v - (uint32_t)dump_log2 <= 8) ? v - (uint32_t)dump_log2 : (u - 2 < K ? u - 2 : K);
    int32_t rc = zb_mle_block_sums(ctx, poly, have, S); // S holds 2^have block sums (blocks of >= 4 entries)
    if (rc) return rc;
    zh_transcript tr; // State.init -> FiatShamirTranscript.init (sumcheck_protocol.zig:149-164)
    zb_mle cur = poly;
    bool owned = false;
    uint32_t round = 0;
    for (;;) {
        uint64_t len = 1ull << have;
        for (uint32_t t = 0; t < have; t++, round++) {
            const uint64_t h = len / 2;
            uint64_t s0 = 0, s1 = 0; // roundPolynomial :216-224 on the block sums
            for (uint64_t i = 0; i < h; i++) {
                s0 = f_add(s0, S[i]);
                s1 = f_add(s1, S[i + h]);
            }
            const uint64_t c[2] = {s0, f_sub(s1, s0)}; // :229
            round_polys[2 * (size_t)round] = c[0];
            round_polys[2 * (size_t)round + 1] = c[1];
            if (round == 0 && claimed_sum) *claimed_sum = f_add(s0, s1); // sumOverHypercube, sumcheck_prover.zig:40
            uint64_t r;
            if (fixed_challenges) {
                r = fixed_challenges[round]; // proveInteractive :127
            } else {
                zh_transcript_append_fields(&tr, c, 2); // generateChallenge, sumcheck_protocol.zig:176-184
                r = zh_transcript_challenge(&tr);
            }
            final_point[round] = r;
            rs[t < 10 ? t : 9] = r;
            for (uint64_t i = 0; i < h; i++) S[i] = f_add(S[i], f_mul(r, f_sub(S[i + h], S[i]))); // partialEval :166-173
            len = h;
        }
        // the device binds the same `have` variables in one pass and returns the sums for the next rounds — or the folded
        // table itself once it is small, on which the host runs the remaining rounds (four lanes wide)
        const uint32_t u_next = u - have;
        const bool publish = u_next <= (uint32_t)dump_log2;
        const uint32_t k_next = publish ? u_next : (u_next - 2 < K ? u_next - 2 : K);
        zb_mle next = 0;
        rc = zb_mle_fold_multi(ctx, cur, have, rs, (consume || owned) ? nullptr : &next, k_next, S);
        if (rc) break;
        if (next) {
            cur = next;
            owned = true;
        }
        u = u_next;
        have = k_next;
        if (publish) {
            const uint64_t m = 1ull << u;
            for (uint64_t i = 0; i < m; i++) T[i] = (uint32_t)S[i];
            finish_rounds_on_host(1, T, m, round, &tr, fixed_challenges, round_polys, final_point, final_eval, nullptr);
            break;
        }
    }
    if (rc == ZB_OK && consume) rc = zb_mle_collapse(ctx, poly, *final_eval); // the caller's table ends as its final evaluation, as after v folds
    if (owned) zb_mle_free(ctx, cur);
    return rc;
}

static int32_t prove_rounds(zb_ctx *ctx, const zb_mle *polys, uint32_t d, bool consume, const uint64_t *fixed_challenges,
                            uint64_t *round_polys, uint64_t *final_point, uint64_t *final_evals, uint64_t *claimed_sum) {
    if (d < 1 || d > 3 || !polys) return ZB_ERR_BAD_ARGUMENT;
    uint64_t n = 0;
    uint32_t v_local = 0;
    int32_t rc = zb_mle_len(ctx, polys[0], &n, &v_local);
    if (rc) return rc;
    int32_t rank = 0, world = 1;
    zb_comm_info(ctx, &rank, &world);
    int64_t lin_dump = 12;
    zb_get_option(ctx, "host_tail_log2", &lin_dump);
    if (d == 1 && world == 1 && v_local > (uint32_t)lin_dump && v_local > 10) {
        int64_t lin = 0;
        zb_get_option(ctx, "linear_d1", &lin);
        if (lin) return prove_linear(ctx, polys[0], v_local, consume, fixed_challenges, round_polys, final_point, final_evals, claimed_sum);
    }
    uint32_t v_tail = 0;
    while ((1 << v_tail) < world) v_tail++;
    const uint32_t v = v_local + v_tail;
    if (v == 0) return ZB_ERR_NO_VARIABLES; // sumcheck_prover.zig:30-32
    const uint32_t nc = d + 1;
    bool sharded = world > 1;
    // while sharded: device-side reduction on, persistent tail off (the per-round collective needs the stream)
    int64_t tail_log2 = 0, comm_reduce = 0, host_tail = 0;
    zb_get_option(ctx, "tail_log2", &tail_log2);
    zb_get_option(ctx, "comm_reduce", &comm_reduce);
    zb_get_option(ctx, "prod_host_tail_log2", &host_tail);
    if (d == 1 && host_tail > 0 && lin_dump > host_tail) host_tail = lin_dump; // one table: its rounds are cheaper still on the host
    struct Restore {
        zb_ctx *c;
        int64_t tail, red;
        bool armed;
        void now() {
            if (armed) {
                zb_set_option(c, "tail_log2", tail);
                zb_set_option(c, "comm_reduce", red);
            }
            armed = false;
        }
        ~Restore() { now(); }
    } restore{ctx, tail_log2, comm_reduce, sharded};
    if (sharded) {
        int64_t p2p = 0;
        zb_get_option(ctx, "p2p_attached", &p2p);
        zb_set_option(ctx, "tail_log2", 0);
        zb_set_option(ctx, "comm_reduce", p2p ? 2 : 1); // NVLink peer exchange inside the kernels when attached, else NCCL
    }
    zh_transcript tr; // State.init -> FiatShamirTranscript.init (sumcheck_protocol.zig:149-164)
    uint64_t coeffs[4];
    uint64_t grid[16];
    alignas(32) uint32_t host_tables[3u << 12]; // tables of <= 2^12 entries finish on the host (zb_prod_fold_dump)
    // tables that are small and whole (not sharded) leave the device: the host finishes their rounds
    auto host_finishes = [&](uint64_t len) { return !sharded && host_tail > 0 && len <= (1ull << host_tail); };
    auto collapse_consumed = [&]() { return consume ? zb_prod_collapse(ctx, polys, d, final_evals) : ZB_OK; };
    if (host_finishes(n)) {
        rc = zb_prod_fold_dump(ctx, polys, d, 0, nullptr, host_tables);
        if (rc) return rc;
        finish_rounds_on_host(d, host_tables, n, 0, &tr, fixed_challenges, round_polys, final_point, final_evals, claimed_sum);
        return collapse_consumed();
    }
    // two rounds per pass over the data (zb_prod_grid / zb_prod_fold_grid) while the folded tables keep >= 2^gmin entries
    const int gmin = grid_min_log2(host_tail > 0);
    const uint64_t grid_min_n = gmin ? (1ull << gmin) : ~0ull;
    auto grid_serves = [&](uint64_t folded_len) { return gmin && folded_len >= grid_min_n && folded_len >= 32; };
    // first pass: "grid" = G of the raw tables (no fold; arithmetic-heavy for d = 3), or "sums" = the plain round-0 sums, after
    // which the first fold already produces the grid of rounds 1 and 2
    static const bool first_is_grid = [] {
        const char *e = getenv("ZB_GRID_FIRST");
        return e && !strcmp(e, "grid");
    }();
    bool have_grid = false; // `grid` holds G of the current tables: this round's AND the next round's polynomial
    if (first_is_grid && gmin && n >= 64) {
        rc = zb_prod_grid(ctx, polys, d, grid);
        if (rc) return rc;
        grid_round_a(d, grid, coeffs);
        have_grid = true;
    } else {
        rc = zb_prod_round_coeffs(ctx, polys, d, coeffs);
        if (rc) return rc;
    }
    if (claimed_sum) {
        // sum over the hypercube == g(0) + g(1) == 2 a0 + a1 + ... + ad   (== sumOverHypercube for d == 1, :40)
        uint64_t s = coeffs[0];
        for (uint32_t k = 0; k < nc; k++) s = f_add(s, coeffs[k]);
        *claimed_sum = s;
    }
    zb_mle cur[3] = {polys[0], d > 1 ? polys[1] : 0, d > 2 ? polys[2] : 0};
    bool owned = false; // `cur` are tables of ours (the caller's stay intact unless `consume`)
    uint64_t n_cur = n; // current (local) table length
    auto cleanup = [&]() {
        if (owned)
            for (uint32_t k = 0; k < d; k++)
                if (cur[k]) zb_mle_free(ctx, cur[k]);
        owned = false;
    };
    // Invariant at the top of the loop: `coeffs` is the polynomial of round `round` over the current tables; with have_grid,
    // `grid` also determines round + 1 once this round's challenge is known. Each pass binds one or two variables.
    uint32_t round = 0;
    while (round < v) {
        if (sharded && n_cur <= (1ull << gather_log2())) {
            // leave the sharded regime: all-gather the shards into the global order; what is in hand (`coeffs`, `grid`) are
            // global sums already, so nothing is recomputed
            zb_mle full[3] = {0, 0, 0};
            rc = zb_comm_allgather_cyclic_batch(ctx, cur, d, full);
            if (rc) {
                cleanup();
                return rc;
            }
            cleanup();
            for (uint32_t k = 0; k < d; k++) cur[k] = full[k];
            owned = true;
            sharded = false;
            n_cur *= (uint64_t)world;
            restore.now();
        }
        const uint32_t nf = have_grid ? 2 : 1;
        uint64_t rr[2] = {0, 0};
        for (uint32_t t = 0; t < nf; t++) {
            if (t == 1) grid_round_b(d, grid, rr[0], coeffs);
            for (uint32_t k = 0; k < nc; k++) round_polys[(size_t)(round + t) * nc + k] = coeffs[k];
            if (fixed_challenges) {
                rr[t] = fixed_challenges[round + t]; // proveInteractive :127
            } else {
                zh_transcript_append_fields(&tr, coeffs, nc); // generateChallenge, sumcheck_protocol.zig:176-184
                rr[t] = zh_transcript_challenge(&tr);
            }
            final_point[round + t] = rr[t];
            // (the reference also evaluates the round polynomial at r to advance its claim, :63-70; the value never
            //  reaches the proof, so it is not computed here)
        }
        round += nf;
        const uint64_t n_after = n_cur >> nf;
        have_grid = false;
        if (host_finishes(n_after)) {
            // the folded tables go to the host, which finishes the proof (the device tables are left as they are)
            rc = zb_prod_fold_dump(ctx, cur, d, nf, rr, host_tables);
            if (rc == ZB_OK)
                finish_rounds_on_host(d, host_tables, n_after, round, &tr, fixed_challenges, round_polys, final_point, final_evals, nullptr);
            const bool callers = !owned;
            cleanup();
            if (rc == ZB_OK && callers) rc = collapse_consumed();
            return rc;
        }
        const bool fresh = !owned && !consume; // the caller's tables must stay intact: fold into new ones
        if (grid_serves(n_after)) {
            zb_mle next[3] = {0, 0, 0};
            rc = zb_prod_fold_grid(ctx, cur, d, nf, rr, fresh ? next : nullptr, grid);
            if (rc) {
                cleanup();
                return rc;
            }
            if (fresh) {
                for (uint32_t k = 0; k < d; k++) cur[k] = next[k];
                owned = true;
            }
            n_cur = n_after;
            grid_round_a(d, grid, coeffs);
            have_grid = true;
            continue;
        }
        // one round per kernel (the persistent tail serves small tables); the last call returns the coefficients of the round
        // after them — or, after the last fold, the d final evaluations (current_poly.evaluations[0], :88)
        if (fresh) {
            zb_mle next[3] = {0, 0, 0};
            rc = zb_prod_partial_eval(ctx, cur, d, rr[0], next, coeffs);
            if (rc) return rc;
            for (uint32_t k = 0; k < d; k++) cur[k] = next[k];
            owned = true;
        } else {
            rc = zb_prod_fold_inplace(ctx, cur, d, rr[0], coeffs);
        }
        if (rc == ZB_OK && nf == 2) rc = zb_prod_fold_inplace(ctx, cur, d, rr[1], coeffs);
        if (rc) {
            cleanup();
            return rc;
        }
        n_cur = n_after;
    }
    for (uint32_t k = 0; k < d; k++) final_evals[k] = coeffs[k];
    cleanup();
    return ZB_OK;
}

// Multi-device context (zb_ctx_create_mask): the tables are sharded cyclically over the GPUs; every rank's host thread
// runs the SAME round loop on its own shard (prove_rounds with world > 1: per-round partial sums exchanged inside the
// kernels over NVLink, identical transcripts on every rank), rank 0 writes the caller's buffers.
namespace {
constexpr uint64_t GROUP_HANDLE_BIT = 1ull << 63;
struct GroupProve {
    zb_ctx *front;
    const zb_mle *polys;
    uint32_t d, v;
    bool consume;
    const uint64_t *fixed;
    uint64_t *round_polys, *final_point, *final_evals, *claimed_sum;
};
int32_t group_prove_rank(zb_ctx *c, int32_t rank, int32_t, void *user) {
    GroupProve *j = static_cast<GroupProve *>(user);
    zb_mle h[3] = {0, 0, 0};
    for (uint32_t k = 0; k < j->d; k++) {
        const int32_t rc = zb_group_mle(j->front, j->polys[k], rank, &h[k]);
        if (rc) return rc;
    }
    if (rank == 0) return prove_rounds(c, h, j->d, j->consume, j->fixed, j->round_polys, j->final_point, j->final_evals, j->claimed_sum);
    std::vector<uint64_t> rp((size_t)j->v * (j->d + 1)), fp(j->v);
    uint64_t fe[3], cs = 0;
    return prove_rounds(c, h, j->d, j->consume, j->fixed, rp.data(), fp.data(), fe, &cs);
}
bool on_group(zb_ctx *ctx, const zb_mle *polys) { return zb_group_size(ctx) > 1 && polys && (polys[0] & GROUP_HANDLE_BIT); }
int32_t group_prove(zb_ctx *ctx, const zb_mle *polys, uint32_t d, bool consume, const uint64_t *fixed, uint64_t *round_polys,
                    uint64_t *final_point, uint64_t *final_evals, uint64_t *claimed_sum) {
    if (d < 1 || d > 3) return ZB_ERR_BAD_ARGUMENT;
    uint64_t n = 0;
    uint32_t v = 0;
    int32_t rc = zb_mle_len(ctx, polys[0], &n, &v);
    if (rc) return rc;
    if (v == 0) return ZB_ERR_NO_VARIABLES;
    for (uint32_t k = 1; k < d; k++) {
        uint64_t nk = 0;
        rc = zb_mle_len(ctx, polys[k], &nk, nullptr);
        if (rc) return rc;
        if (nk != n) return ZB_ERR_DIFFERENT_NUM_VARS;
    }
    GroupProve job{ctx, polys, d, v, consume, fixed, round_polys, final_point, final_evals, claimed_sum};
    rc = zb_group_run(ctx, group_prove_rank, &job);
    if (rc == ZB_OK && consume) // the shards were folded away in place; what is left of them is not a table of the caller's any more
        for (uint32_t k = 0; k < d; k++) zb_group_mle_set_len(ctx, polys[k], (uint64_t)zb_group_size(ctx));
    return rc;
}
} // namespace

// measurement helper (like zb_int_pipe_peak / zb_h2d_rate): `reps` proves of `poly` in a row through this very ABI, wall-clock
// microseconds per prove — the Python mirror's per-call overhead (~10 us) would otherwise be part of a 50 us measurement
int32_t zh_time_sumcheck_prove(zb_ctx *ctx, zb_mle poly, uint32_t reps, double *us_per_prove) {
    uint64_t n = 0;
    uint32_t v = 0;
    int32_t rc = zb_mle_len(ctx, poly, &n, &v);
    if (rc || !us_per_prove || reps == 0) return rc ? rc : ZB_ERR_BAD_ARGUMENT;
    std::vector<uint64_t> rp(2 * (size_t)(v ? v : 1)), fp(v ? v : 1);
    uint64_t fe = 0, cs = 0;
    for (int i = 0; i < 3 && rc == ZB_OK; i++) rc = zh_sumcheck_prove(ctx, poly, rp.data(), fp.data(), &fe, &cs);
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < reps && rc == ZB_OK; i++) rc = zh_sumcheck_prove(ctx, poly, rp.data(), fp.data(), &fe, &cs);
    *us_per_prove = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    return rc;
}

int32_t zh_prodcheck_finish_small(uint32_t d, uint32_t *tables, uint64_t m, uint32_t round, zh_transcript *tr,
                                  const uint64_t *fixed_challenges, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_evals) {
    if (d < 1 || d > 3 || !tables || m < 1 || m > 4096 || (m & (m - 1)) || (!tr && !fixed_challenges) || !final_evals) return ZB_ERR_BAD_ARGUMENT;
    if (m > 1 && (!round_polys || !final_point)) return ZB_ERR_BAD_ARGUMENT;
    for (uint64_t i = 0; i < (uint64_t)d * m; i++)
        if (tables[i] >= P) return ZB_ERR_NOT_CANONICAL;
    finish_rounds_on_host(d, tables, m, round, tr, fixed_challenges, round_polys, final_point, final_evals, nullptr);
    return ZB_OK;
}

int32_t zh_set_grid_min_log2(int32_t v) {
    return g_grid_min_log2.exchange(v < 0 ? -1 : v, std::memory_order_relaxed); // -1 = back to the default
}

int32_t zh_sumcheck_prove(zb_ctx *ctx, zb_mle poly, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval,
                          uint64_t *claimed_sum) {
    if (on_group(ctx, &poly)) return group_prove(ctx, &poly, 1, false, nullptr, round_polys, final_point, final_eval, claimed_sum);
    return prove_rounds(ctx, &poly, 1, false, nullptr, round_polys, final_point, final_eval, claimed_sum);
}

int32_t zh_sumcheck_prove_interactive(zb_ctx *ctx, zb_mle poly, const uint64_t *challenges, uint32_t n_challenges,
                                      uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval) {
    uint64_t n;
    uint32_t v;
    int32_t rc = zb_mle_len(ctx, poly, &n, &v);
    if (rc) return rc;
    if (v == 0) return ZB_ERR_NO_VARIABLES;                       // :102-104
    if (n_challenges != v) return ZB_ERR_WRONG_NUM_CHALLENGES;    // :105-107
    for (uint32_t i = 0; i < v; i++)
        if (challenges[i] >= P) return ZB_ERR_NOT_CANONICAL;
    if (on_group(ctx, &poly)) return group_prove(ctx, &poly, 1, false, challenges, round_polys, final_point, final_eval, nullptr);
    return prove_rounds(ctx, &poly, 1, false, challenges, round_polys, final_point, final_eval, nullptr);
}

size_t zh_sumcheck_proof_to_bytes(uint32_t v, const uint64_t *round_polys, const uint64_t *final_point, uint64_t final_eval,
                                  uint8_t *out) { // sumcheck_protocol.zig:76-109, little-endian u64s
    uint64_t *w = reinterpret_cast<uint64_t *>(out);
    size_t k = 0;
    auto put = [&](uint64_t x) { memcpy(out + 8 * k++, &x, 8); };
    (void)w;
    put(v);
    for (uint32_t i = 0; i < 2 * v; i++) put(round_polys[i]);
    for (uint32_t i = 0; i < v; i++) put(final_point[i]);
    put(final_eval);
    return 8 * k;
}

int32_t zh_prodcheck_prove(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *round_polys, uint64_t *final_point,
                           uint64_t *final_evals, uint64_t *claimed_sum) {
    if (on_group(ctx, polys)) return group_prove(ctx, polys, d, false, nullptr, round_polys, final_point, final_evals, claimed_sum);
    return prove_rounds(ctx, polys, d, false, nullptr, round_polys, final_point, final_evals, claimed_sum);
}

int32_t zh_prodcheck_prove_consume(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *round_polys, uint64_t *final_point,
                                   uint64_t *final_evals, uint64_t *claimed_sum) {
    if (on_group(ctx, polys)) return group_prove(ctx, polys, d, true, nullptr, round_polys, final_point, final_evals, claimed_sum);
    return prove_rounds(ctx, polys, d, true, nullptr, round_polys, final_point, final_evals, claimed_sum);
}

int32_t zh_eqcheck_prove(zb_ctx *ctx, const uint64_t *tau, uint32_t num_vars, const zb_mle *polys, uint32_t d, uint64_t *round_polys,
                         uint64_t *final_point, uint64_t *final_evals, uint64_t *claimed_sum) {
    if (d < 1 || d > 2 || !polys || (num_vars && !tau)) return ZB_ERR_BAD_ARGUMENT;
    if (zb_group_size(ctx) > 1 && (polys[0] >> 63)) return ZB_ERR_BAD_ARGUMENT; // sharded tables: not built for this extension
    uint64_t n = 0;
    uint32_t v = 0;
    int32_t rc = zb_mle_len(ctx, polys[0], &n, &v);
    if (rc) return rc;
    if (v != num_vars) return ZB_ERR_WRONG_NUM_VARS;
    if (v == 0) return ZB_ERR_NO_VARIABLES;
    zb_mle all[3] = {0, polys[0], d > 1 ? polys[1] : 0};
    rc = zb_mle_eq(ctx, tau, v, &all[0]);
    if (rc) return rc;
    rc = prove_rounds(ctx, all, d + 1, false, nullptr, round_polys, final_point, final_evals, claimed_sum);
    zb_mle_free(ctx, all[0]);
    return rc;
}

/* ------------------------------------------------------------------ commitment scheme */

int32_t zh_commit(zb_ctx *ctx, zb_mle poly, zb_tree *tree, uint8_t root[32], uint32_t *num_vars) { // :69-83
    uint64_t n;
    uint32_t v;
    int32_t rc = zb_mle_len(ctx, poly, &n, &v);
    if (rc) return rc;
    rc = zb_merkle_build(ctx, &poly, 1, tree, root);
    if (rc == ZB_OK && num_vars) *num_vars = v;
    return rc;
}

int32_t zh_batch_commit(zb_ctx *ctx, const zb_mle *polys, uint32_t count, zb_tree *trees, uint8_t *roots) { // :132-157
    if (count == 0) return ZB_OK;
    return zb_merkle_build(ctx, polys, count, trees, roots);
}

int32_t zh_commit_sharded(zb_ctx *ctx, zb_mle local_poly, zb_tree *tree, uint8_t local_root[32], uint8_t root[32]) {
    // Subtree sharding: rank g holds the CONTIGUOUS leaves [g N/P, (g+1) N/P); its subtree root is
    // the node at height log2(N/P); the top log2(P) levels are P-1 host hashes (mergeHashesSHA3, hash.zig:187-195).
    int32_t rank = 0, world = 1;
    zb_comm_info(ctx, &rank, &world);
    if (world > 16) return ZB_ERR_BAD_ARGUMENT;
    uint8_t mine[32];
    int32_t rc = zb_merkle_build(ctx, &local_poly, 1, tree, mine);
    if (rc) return rc;
    if (local_root) memcpy(local_root, mine, 32);
    uint64_t slots[64] = {0};
    memcpy(slots + 4 * rank, mine, 32);
    rc = zb_comm_allreduce_u64(ctx, slots, 4 * (uint32_t)world); // all-gather as a sum with zeros: exact
    if (rc) return rc;
    uint8_t level[16][32];
    memcpy(level, slots, 32 * (size_t)world);
    for (int32_t w = world; w > 1; w /= 2)
        for (int32_t i = 0; i < w / 2; i++) {
            uint8_t buf[64];
            memcpy(buf, level[2 * i], 32);
            memcpy(buf + 32, level[2 * i + 1], 32);
            Sha3_256::hash(buf, 64, level[i]);
        }
    memcpy(root, level[0], 32);
    return ZB_OK;
}

uint64_t zh_point_to_index(const uint64_t *point, uint32_t npoint) { // :178-183
    if (npoint == 0) return 0;
    return npoint >= 64 ? point[0] : point[0] % (1ull << npoint);
}

int32_t zh_commit_open(zb_ctx *ctx, zb_mle poly, zb_tree tree, const uint64_t *point, uint32_t npoint, uint64_t *value,
                       uint64_t *leaf_index, uint64_t *leaf_value, uint8_t *siblings, uint8_t *dirs) { // :86-115
    uint64_t n;
    uint32_t v;
    int32_t rc = zb_mle_len(ctx, poly, &n, &v);
    if (rc) return rc;
    if (npoint != v) return ZB_ERR_POINT_DIM_MISMATCH; // :92-94
    rc = zb_mle_eval(ctx, poly, point, npoint, value);  // :97
    if (rc) return rc;
    uint64_t index = zh_point_to_index(point, npoint); // :102
    if (leaf_index) *leaf_index = index;
    return zb_merkle_open(ctx, tree, index, siblings, dirs, leaf_value); // :105
}

int32_t zh_merkle_verify(const uint8_t root[32], uint64_t value, const uint8_t *siblings, const uint8_t *dirs, uint32_t height) {
    // merkle_tree.zig:362-373: current = hashLeaf(value); per level current = hashInternal(left, right)
    uint8_t cur[32], buf[64];
    Sha3_256::hash(&value, 8, cur);
    for (uint32_t l = 0; l < height; l++) {
        if (dirs[l]) { // we are the right child
            memcpy(buf, siblings + 32 * l, 32);
            memcpy(buf + 32, cur, 32);
        } else {
            memcpy(buf, cur, 32);
            memcpy(buf + 32, siblings + 32 * l, 32);
        }
        Sha3_256::hash(buf, 64, cur);
    }
    return memcmp(cur, root, 32) == 0;
}

int32_t zh_commit_verify(const uint8_t root[32], uint64_t leaf_value, const uint8_t *siblings, const uint8_t *dirs,
                         uint32_t height) { // polynomial_commit.zig:118-129 (dimension check is the caller's: height)
    return zh_merkle_verify(root, leaf_value, siblings, dirs, height);
}

static int32_t commitments_after_build(zb_ctx *ctx, zh_transcript *tr, const zb_mle *polys, std::vector<zb_tree> &trees, uint32_t count,
                                       uint8_t *roots, uint64_t *points, uint64_t *values, uint64_t *leaf_indices, uint64_t *leaf_values,
                                       uint8_t *siblings, uint8_t *dirs);

int32_t zh_generate_commitments(zb_ctx *ctx, zh_transcript *tr, const zb_mle *polys, uint32_t count, uint8_t *roots,
                                uint64_t *points, uint64_t *values, uint64_t *leaf_indices, uint64_t *leaf_values,
                                uint8_t *siblings, uint8_t *dirs) {
    // Prover.generateCommitments, src/prover/prover.zig:366-467. Same transcript traffic, same outputs; the 43 commits
    // are ONE batched device build (trees retained), each opening is an O(N) evaluation + a gather.
    if (!tr || !polys || count == 0) return ZB_ERR_BAD_ARGUMENT;
    std::vector<zb_tree> trees(count, 0);
    const int32_t rc = zb_merkle_build(ctx, polys, count, trees.data(), roots); // PHASE 1 (:405-410)
    if (rc) return rc;
    return commitments_after_build(ctx, tr, polys, trees, count, roots, points, values, leaf_indices, leaf_values, siblings, dirs);
}

// PHASES 2-4 of Prover.generateCommitments on trees that exist already (built by zb_merkle_build, or by the pipelined
// zb_witness_pack_commit of zh_prove_from_trace); the trees are released on return (:446-448)
static int32_t commitments_after_build(zb_ctx *ctx, zh_transcript *tr, const zb_mle *polys, std::vector<zb_tree> &trees, uint32_t count,
                                       uint8_t *roots, uint64_t *points, uint64_t *values, uint64_t *leaf_indices, uint64_t *leaf_values,
                                       uint8_t *siblings, uint8_t *dirs) {
    uint64_t n;
    uint32_t v;
    int32_t rc = zb_mle_len(ctx, polys[0], &n, &v);
    if (rc) {
        for (uint32_t i = 0; i < count; i++) zb_merkle_free(ctx, trees[i]);
        return rc;
    }
    zh_transcript_append_bytes(tr, "POLY_COMMITMENTS", 16); // PHASE 2 (:413-416)
    for (uint32_t i = 0; i < count; i++) zh_transcript_append_bytes(tr, roots + 32 * (size_t)i, 32);
    // PHASE 3 (:420-443). The loop only draws challenges from the transcript (nothing is absorbed before PHASE 4), so
    // all opening points exist before the first value is needed: the `count` evaluations are one batched launch set and
    // the `count` Merkle paths one gather, two read-backs in total instead of 2 * count.
    for (uint32_t i = 0; i < count; i++)
        for (uint32_t j = 0; j < v; j++) points[(size_t)i * v + j] = zh_transcript_challenge(tr);
    if (zb_group_size(ctx) > 1 && (polys[0] >> 63)) { // sharded tables of a multi-device context: one opening at a time
        for (uint32_t i = 0; i < count && rc == ZB_OK; i++)
            rc = zh_commit_open(ctx, polys[i], trees[i], points + (size_t)i * v, v, &values[i], &leaf_indices[i], &leaf_values[i],
                                siblings + (size_t)i * v * 32, dirs + (size_t)i * v);
    } else {
        // :427 and Scheme.open :431 evaluate the same polynomial at the same point twice; once is enough
        rc = zb_mle_eval_batch(ctx, polys, count, points, v, values);
        for (uint32_t i = 0; i < count; i++) leaf_indices[i] = zh_point_to_index(points + (size_t)i * v, v); // :102
        if (rc == ZB_OK) rc = zb_merkle_open_batch(ctx, trees.data(), count, leaf_indices, siblings, dirs, leaf_values);
    }
    for (uint32_t i = 0; i < count; i++) zb_merkle_free(ctx, trees[i]); // :446-448
    if (rc) return rc;
    zh_transcript_append_bytes(tr, "OPENING_CLAIMS", 14); // PHASE 4 (:463-466)
    zh_transcript_append_fields(tr, values, count);
    return ZB_OK;
}

/* ------------------------------------------------------------------ prove (after the VM), serialize, verify */

namespace {
// SHA-256 (FIPS 180-4) for the program hash (std.crypto.hash.sha2.Sha256 in the reference, prover.zig:98-99)
struct Sha256 {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t *p) {
        static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
            0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
            0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
            0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
            0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
            0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    static void hash(const void *data, size_t len, uint8_t out[32]) {
        Sha256 s;
        const uint8_t *p = static_cast<const uint8_t *>(data);
        size_t n = len;
        for (; n >= 64; n -= 64, p += 64) s.block(p);
        uint8_t blk[128] = {0};
        memcpy(blk, p, n);
        blk[n] = 0x80;
        const size_t tot = n + 9 <= 64 ? 64 : 128;
        const uint64_t bits = (uint64_t)len * 8;
        for (int i = 0; i < 8; i++) blk[tot - 1 - i] = (uint8_t)(bits >> (8 * i));
        s.block(blk);
        if (tot == 128) s.block(blk + 64);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(s.h[i] >> 24); out[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
            out[4 * i + 2] = (uint8_t)(s.h[i] >> 8); out[4 * i + 3] = (uint8_t)s.h[i];
        }
    }
};

struct ByteWriter { // little-endian writer over a caller buffer (std.io.fixedBufferStream in the reference); never writes past `end`
    uint8_t *p, *end;
    bool overflow = false;
    void bytes(const void *d, size_t n) {
        if ((size_t)(end - p) < n) {
            overflow = true;
            return;
        }
        memcpy(p, d, n);
        p += n;
    }
    void u8(uint8_t v) { bytes(&v, 1); }
    void u32(uint32_t v) { bytes(&v, 4); }
    void u64(uint64_t v) { bytes(&v, 8); }
};
struct ByteReader {
    const uint8_t *p, *end;
    bool need(size_t n) const { return (size_t)(end - p) >= n; }
    uint32_t u32() { uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
    uint64_t u64() { uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
};

uint32_t log2_ceil_u64(uint64_t n) { // std.math.log2_int_ceil
    uint32_t v = 0;
    while ((1ull << v) < n) v++;
    return v;
}
bool opcode_has_table(uint64_t op) { // getTableMetadata, src/isa/instruction_table.zig:243-275 (OP, OP_IMM, LOAD, STORE, BRANCH)
    return op == 0x33 || op == 0x13 || op == 0x03 || op == 0x23 || op == 0x63;
}
} // namespace

void zh_sha256(const void *data, size_t n, uint8_t out[32]) { Sha256::hash(data, n, out); }

int32_t zh_prove_from_trace(zb_ctx *ctx, const uint8_t *program, size_t program_len, uint64_t entry_pc,
                            const uint64_t *initial_regs, uint32_t n_init, const uint64_t *cols, uint64_t num_steps, uint64_t final_pc,
                            const uint64_t *final_regs, const uint64_t *outputs, uint32_t n_out, int32_t compat_buffer, uint8_t *out,
                            size_t out_cap, size_t *out_len) {
    if (num_steps == 0) return ZB_ERR_EMPTY_TRACE; // prover.zig:144-146
    if (!cols || !final_regs || (n_init && !initial_regs) || (n_out && !outputs) || (program_len && !program)) return ZB_ERR_BAD_ARGUMENT;
    const uint32_t v = log2_ceil_u64(num_steps); // Proof.init, proof.zig:225
    uint64_t n_lookups = 0; // ConstraintSystem.extractLookupConstraints, builder.zig:253-267
    for (uint64_t i = 0; i < num_steps; i++) n_lookups += opcode_has_table(cols[33 * num_steps + i]);
    const size_t exact = 32 + (32 + 8 + 8 + 4 + 8 * (size_t)n_init + 4 + 8 * 32 + 8 + 4 + 8 * (size_t)n_out) + ((size_t)v * 40 + 8) + 4 +
                         (size_t)n_lookups * 24 + 43 * (68 + (size_t)v * 41);
    if (compat_buffer) { // estimateSize, serialization.zig:134-173
        const size_t est = 32 + (32 + 8 + 8 + 4 + 4) + 8 * (size_t)n_init + 8 * 32 + ((size_t)v * 40 + 8) + 4 + (size_t)n_lookups * 20 +
                           43 * (32 + (size_t)v * 8 + 8 + 640);
        if (exact > est) return ZB_ERR_NO_SPACE_LEFT;
    }
    if (out_len) *out_len = exact;
    if (!out || out_cap < exact) return ZB_ERR_OOM;
    zh_transcript tr; // prover.zig:91
    uint8_t program_hash[32];
    Sha256::hash(program, program_len, program_hash); // :98-100
    zh_transcript_append_bytes(&tr, program_hash, 32);
    zh_transcript_append_field(&tr, entry_pc % P); // :103
    for (uint32_t i = 0; i < n_init; i++) zh_transcript_append_field(&tr, initial_regs[i] % P); // :106-110
    // witness polynomials straight into HBM (witness.zig:29-270)
    // ... and their commitments in the same pipeline (the trace upload overlaps the leaf hashing): PHASE 1 of
    // Prover.generateCommitments (prover.zig:405-410) needs nothing from the transcript, only the polynomials
    zb_mle polys[43] = {0};
    std::vector<zb_tree> trees(43, 0);
    std::vector<uint8_t> roots(43 * 32);
    uint32_t nv = 0;
    int32_t rc;
    if (num_steps >= (1u << 15)) {
        rc = zb_witness_pack_commit(ctx, cols, num_steps, 43, 33, polys, &nv, trees.data(), roots.data());
    } else { // short traces: 3 x 43 tiny launches cost more than the overlap saves (1.9 vs 1.45 ms at 2^12 steps)
        rc = zb_witness_pack(ctx, cols, num_steps, 43, 33, polys, &nv);
        if (rc == ZB_OK) {
            rc = zb_merkle_build(ctx, polys, 43, trees.data(), roots.data());
            if (rc)
                for (int i = 0; i < 43; i++) zb_mle_free(ctx, polys[i]);
        }
    }
    if (rc) return rc;
    ByteWriter w{out, out + exact};
    w.bytes("ZIGZ", 4); // header, serialization.zig:175-182
    w.u32(1); w.u64(P); w.u64(num_steps); w.u32(v); w.u32(0);
    w.bytes(program_hash, 32); // public I/O, :209-245 (packagePublicIO prover.zig:514-559)
    w.u64(entry_pc); w.u64(final_pc);
    w.u32(n_init);
    for (uint32_t i = 0; i < n_init; i++) w.u64(initial_regs[i]);
    w.u32(32);
    for (int i = 0; i < 32; i++) w.u64(final_regs[i]);
    w.u64(num_steps);
    w.u32(n_out);
    for (uint32_t i = 0; i < n_out; i++) w.u64(outputs[i]);
    // constraint sumcheck placeholder: zero round polynomials, transcript challenges (prover.zig:229-289)
    zh_transcript_append_bytes(&tr, "SUMCHECK_BEGIN", 14);
    zh_transcript_append_field(&tr, num_steps % P);
    zh_transcript_append_field(&tr, v);
    uint64_t chal[64];
    const uint64_t zeros[4] = {0, 0, 0, 0};
    for (uint32_t r = 0; r < v; r++) {
        zh_transcript_append_fields(&tr, zeros, 4);
        chal[r] = zh_transcript_challenge(&tr);
    }
    for (uint32_t r = 0; r < 4 * v; r++) w.u64(0); // serialization.zig:296-311
    for (uint32_t r = 0; r < v; r++) w.u64(chal[r]);
    w.u64(0);
    // Lasso placeholders: one 0-round proof per lookup constraint (prover.zig:292-362, serialization.zig:333-344)
    zh_transcript_append_bytes(&tr, "LASSO_BEGIN", 11);
    w.u32((uint32_t)n_lookups);
    for (uint64_t k = 0; k < n_lookups; k++) {
        zh_transcript_append_bytes(&tr, "LASSO_TABLE", 11);
        zh_transcript_append_field(&tr, (uint32_t)k % P);
        w.u32((uint32_t)k); w.u64(1); w.u32(0); w.u64(0);
    }
    // commitments + openings on the device (prover.zig:366-467)
    const size_t vv = v ? v : 1;
    std::vector<uint8_t> sib(43 * vv * 32), dirs(43 * vv);
    std::vector<uint64_t> pts(43 * vv), vals(43), li(43), lv(43);
    rc = commitments_after_build(ctx, &tr, polys, trees, 43, roots.data(), pts.data(), vals.data(), li.data(), lv.data(), sib.data(),
                                 dirs.data());
    for (int i = 0; i < 43; i++) zb_mle_free(ctx, polys[i]);
    if (rc) return rc;
    for (int i = 0; i < 43; i++) { // serialization.zig:374-429
        w.bytes(roots.data() + 32 * i, 32);
        for (uint32_t j = 0; j < v; j++) w.u64(pts[(size_t)i * v + j]);
        w.u64(vals[i]);
        w.u64(vals[i]); // OpeningProof.value: Scheme.open evaluates the same polynomial at the same point (polynomial_commit.zig:97)
        w.u64(li[i]); w.u64(lv[i]); w.u32(v);
        w.bytes(sib.data() + (size_t)i * v * 32, (size_t)v * 32);
        for (uint32_t j = 0; j < v; j++) w.u8(dirs[(size_t)i * v + j] ? 1 : 0);
    }
    return (!w.overflow && (size_t)(w.p - out) == exact) ? ZB_OK : ZB_ERR_BAD_ARGUMENT;
}

int32_t zh_verify_proof(const uint8_t *proof, size_t len, const uint8_t *program, size_t program_len, int32_t *verdict) {
    if (!proof || !verdict) return ZB_ERR_BAD_ARGUMENT;
    ByteReader r{proof, proof + len};
    if (!r.need(32) || memcmp(r.p, "ZIGZ", 4)) return ZB_ERR_INVALID_PROOF; // readHeader, serialization.zig:184-207
    r.p += 4;
    if (r.u32() != 1) return ZB_ERR_INVALID_PROOF;
    if (r.u64() != P) return ZB_ERR_INVALID_PROOF; // FieldMismatch :108-110
    const uint64_t num_steps = r.u64();
    const uint32_t v = r.u32();
    r.u32();
    if (v > 63 || v != log2_ceil_u64(num_steps)) return ZB_ERR_INVALID_PROOF; // Proof.init derives num_vars from num_steps (:113)
    if (!r.need(52)) return ZB_ERR_INVALID_PROOF;
    const uint8_t *ph = r.p;
    r.p += 32;
    r.u64(); r.u64();
    for (int part = 0; part < 2; part++) { // initial / final registers
        const uint32_t n = r.u32();
        if (!r.need(8 * (size_t)n + 12)) return ZB_ERR_INVALID_PROOF;
        r.p += 8 * (size_t)n;
    }
    r.u64();
    const uint32_t n_out = r.u32();
    if (!r.need(8 * (size_t)n_out)) return ZB_ERR_INVALID_PROOF;
    r.p += 8 * (size_t)n_out;
    // The reference deserializes the WHOLE proof before it verifies anything (serialization.zig:98-127), so a truncated or
    // garbled later section is a deserialize error even when an earlier check would reject: walk every section first and
    // decide afterwards, in the verifier's order (program hash, constraint sumcheck, Lasso proofs, openings).
    *verdict = 0;
    // verifySumcheckProof: only round 0 is checked, g(0) + g(1) == final_eval (verifier.zig:196-214)
    auto sumcheck = [&](uint32_t nv, int ncoef, bool *ok) -> bool {
        if (!r.need((size_t)nv * (ncoef + 1) * 8 + 8)) return false;
        uint64_t g0 = 0, g1 = 0;
        for (uint32_t rd = 0; rd < nv; rd++)
            for (int k = 0; k < ncoef; k++) {
                const uint64_t c = r.u64() % P;
                if (rd == 0) {
                    if (k == 0) g0 = c;
                    g1 = f_add(g1, c);
                }
            }
        r.p += (size_t)nv * 8;
        const uint64_t fe = r.u64() % P;
        *ok = nv == 0 || f_add(g0, g1) == fe;
        return true;
    };
    bool ok = true;
    if (!sumcheck(v, 4, &ok)) return ZB_ERR_INVALID_PROOF;
    if (!ok) *verdict = 1;
    if (!r.need(4)) return ZB_ERR_INVALID_PROOF;
    const uint32_t n_lasso = r.u32();
    for (uint32_t k = 0; k < n_lasso; k++) { // verifyLassoProof :233-262
        if (!r.need(16)) return ZB_ERR_INVALID_PROOF;
        r.u32(); r.u64();
        const uint32_t lv = r.u32();
        if (!sumcheck(lv, 3, &ok)) return ZB_ERR_INVALID_PROOF;
        if (!ok && *verdict == 0) *verdict = 2;
    }
    for (int i = 0; i < 43; i++) { // verifyOpening :270-294
        if (!r.need(32 + (size_t)v * 8 + 36)) return ZB_ERR_INVALID_PROOF;
        const uint8_t *root = r.p;
        r.p += 32 + (size_t)v * 8;
        const uint64_t value = r.u64() % P, pvalue = r.u64() % P;
        r.u64();
        const uint64_t leaf = r.u64() % P;
        const uint32_t plen = r.u32();
        if (!r.need((size_t)plen * 33)) return ZB_ERR_INVALID_PROOF;
        const uint8_t *sibs = r.p, *dirs = r.p + (size_t)plen * 32;
        r.p += (size_t)plen * 33;
        if (*verdict == 0 && (value != pvalue || !zh_merkle_verify(root, leaf, sibs, dirs, plen))) *verdict = 3;
    }
    uint8_t hash[32];
    Sha256::hash(program, program_len, hash);
    if (memcmp(hash, ph, 32)) { // bindPublicInputs, verifier.zig:101-107: an error, raised before any verdict
        *verdict = 0;
        return ZB_ERR_PROGRAM_HASH_MISMATCH;
    }
    return ZB_OK;
}

/* ------------------------------------------------------------------ Lasso */

void zh_flat_commit(const uint64_t *evals, uint64_t n, uint8_t out[32]) { // lasso_prover.zig:242-252
    Sha3_256 h;
    h.update_words(evals, n);
    h.peek(out);
}

void zh_flat_commit_u32(const uint32_t *evals, uint64_t n, uint8_t out[32]) {
    Sha3_256 h;
    h.update_words_u32(evals, n);
    h.peek(out);
}

int32_t zh_lasso_commit_poly(zb_ctx *ctx, zb_mle poly, uint8_t out[32]) { // lasso_prover.zig:242-252
    uint64_t n;
    uint32_t v;
    int32_t rc = zb_mle_len(ctx, poly, &n, &v);
    if (rc) return rc;
    // The sponge is one sequential chain over 8n bytes (n/17 dependent permutations): it stays on one host core.
    // The evaluations come down in their 4-byte device form into the context's pinned scratch and are absorbed as
    // le64 words (zero-extended on the fly).
    const uint64_t CH = 4ull << 20;
    uint32_t *buf = nullptr;
    rc = zb_host_scratch(ctx, (n < CH ? n : CH) * sizeof(uint32_t), (void **)&buf);
    if (rc) return rc;
    Sha3_256 h;
    for (uint64_t off = 0; off < n; off += CH) {
        uint64_t k = n - off < CH ? n - off : CH;
        rc = zb_mle_download_u32(ctx, poly, off, buf, k);
        if (rc) return rc;
        h.update_words_u32(buf, k);
    }
    h.peek(out);
    return ZB_OK;
}

static int32_t lasso_finish(zb_ctx *ctx, zb_mle table_poly, zb_mle query_poly, uint64_t *round_polys, uint64_t *final_point,
                            uint64_t *final_eval, uint32_t *num_vars, uint8_t qc[32], uint8_t tc[32]) {
    uint64_t n;
    uint32_t v;
    int32_t rc = zb_mle_len(ctx, query_poly, &n, &v);
    if (rc) return rc;
    if (num_vars) *num_vars = v;
    // :154 `_ = query_poly.sumOverHypercube()` is discarded by the reference; nothing to do.
    rc = zh_sumcheck_prove(ctx, query_poly, round_polys, final_point, final_eval, nullptr); // :160
    if (rc) return rc;
    rc = zh_lasso_commit_poly(ctx, query_poly, qc); // :163
    if (rc) return rc;
    return zh_lasso_commit_poly(ctx, table_poly, tc); // :164
}

static uint64_t ceil_pow2(uint64_t n) {
    uint64_t r = 1;
    while (r < n) r <<= 1;
    return r;
}

// The sumcheck proof and the two commitments of a Lasso proof do not depend on one another (lasso_prover.zig:160-164),
// and the query commitment — one sequential SHA3 sponge over every padded query evaluation, :242-252 — is by far the
// longest step. For long query lists it therefore runs on a second host thread over a pinned mirror of the query
// polynomial that fills chunk by chunk (zb_xxh3_rows_stream) while this thread uploads the remaining rows, proves the
// sumcheck and commits to the table. Same digests, same proof; zh_set_lasso_pipeline_min_log2(-1) keeps it on one thread.
static std::atomic<int> g_lasso_pipeline_min_log2{-2}; // -2: not read yet, -1: never
static int lasso_pipeline_min_log2() {
    int v = g_lasso_pipeline_min_log2.load(std::memory_order_relaxed);
    if (v == -2) {
        const char *e = getenv("ZB_LASSO_PIPELINE_MIN_LOG2");
        v = e && *e ? atoi(e) : 18;
        if (v < -1) v = -1;
        g_lasso_pipeline_min_log2.store(v, std::memory_order_relaxed);
    }
    return v;
}
static bool lasso_pipeline(uint64_t n_padded) {
    const int v = lasso_pipeline_min_log2();
    return v >= 0 && v < 63 && n_padded >= (1ull << v);
}

// The sponge thread of one query commitment: absorbs mirror[0, n) as le64 words while `avail` grows.
struct SpongeFeed {
    const uint32_t *mirror = nullptr;
    uint64_t n = 0;
    uint64_t avail = 0; // written by the producer (release), read here (acquire)
    int stop = 0;
    Sha3_256 h;
    std::thread th;
    // false when no thread could be started (the caller then commits sequentially)
    bool start(const uint32_t *m, uint64_t count) {
        mirror = m;
        n = count;
        try {
            th = std::thread([this] { run(); });
        } catch (...) {
            return false;
        }
        return true;
    }
    void run() {
        {
            uint64_t done = 0;
            while (done < n) {
                const uint64_t a = __atomic_load_n(&avail, __ATOMIC_ACQUIRE);
                if (a == done) {
                    if (__atomic_load_n(&stop, __ATOMIC_ACQUIRE)) return;
                    _mm_pause();
                    continue;
                }
                h.update_words_u32(mirror + done, a - done);
                done = a;
            }
        }
    }
    void finish(bool ok, uint8_t out[32]) {
        if (!ok) __atomic_store_n(&stop, 1, __ATOMIC_RELEASE);
        if (th.joinable()) th.join();
        if (ok) h.peek(out);
    }
};

// the calling thread's share of a pipelined proof; `feed` has been started on `mirror` and is finished by the caller
static int32_t lasso_piped_body(zb_ctx *ctx, zb_mle table_poly, const uint64_t *query_rows, uint64_t n_queries, uint32_t arity,
                                uint64_t n_padded, uint32_t *mirror, SpongeFeed &feed, uint64_t *round_polys,
                                uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars, uint8_t tc[32]) {
    zb_mle query_poly = 0;
    int32_t rc = zb_xxh3_rows_stream(ctx, query_rows, n_queries, arity, n_padded, &query_poly, mirror, &feed.avail); // :131-142
    if (rc == ZB_OK) {
        if (num_vars) zb_mle_len(ctx, query_poly, nullptr, num_vars);
        rc = zh_sumcheck_prove(ctx, query_poly, round_polys, final_point, final_eval, nullptr); // :160
    }
    if (rc == ZB_OK) rc = zh_lasso_commit_poly(ctx, table_poly, tc); // :164
    if (query_poly) zb_mle_free(ctx, query_poly);
    return rc;
}

// query rows -> query polynomial, sumcheck, both commitments (table_poly is consumed: freed on return)
static int32_t lasso_run(zb_ctx *ctx, zb_mle table_poly, const uint64_t *query_rows, uint64_t n_queries, uint32_t arity,
                         uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars, uint8_t qc[32],
                         uint8_t tc[32]) {
    const uint64_t n_padded = ceil_pow2(n_queries);
    int32_t rc;
    if (!lasso_pipeline(n_padded)) {
        zb_mle query_poly = 0;
        rc = zb_xxh3_rows(ctx, query_rows, n_queries, arity, n_padded, &query_poly); // :131-142
        if (rc == ZB_OK) rc = lasso_finish(ctx, table_poly, query_poly, round_polys, final_point, final_eval, num_vars, qc, tc);
        if (query_poly) zb_mle_free(ctx, query_poly);
    } else {
        uint32_t *mirror = nullptr;
        rc = zb_host_mirror(ctx, n_padded * sizeof(uint32_t), (void **)&mirror);
        if (rc == ZB_OK) {
            SpongeFeed feed;
            const bool threaded = feed.start(mirror, n_padded);
            rc = lasso_piped_body(ctx, table_poly, query_rows, n_queries, arity, n_padded, mirror, feed, round_polys, final_point,
                                  final_eval, num_vars, tc);
            if (!threaded && rc == ZB_OK) feed.run(); // the mirror is complete: absorb it here
            feed.finish(rc == ZB_OK, qc);             // :163
        }
    }
    zb_mle_free(ctx, table_poly);
    return rc;
}

int32_t zh_set_lasso_pipeline_min_log2(int32_t v) {
    const int32_t old = lasso_pipeline_min_log2();
    g_lasso_pipeline_min_log2.store(v < -1 ? -1 : v, std::memory_order_relaxed);
    return old;
}

int32_t zh_lasso_prove(zb_ctx *ctx, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows,
                       uint64_t n_queries, uint32_t arity, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval,
                       uint32_t *num_vars, uint8_t qc[32], uint8_t tc[32]) {
    if (n_queries == 0) return ZB_ERR_NO_QUERIES; // :108-110
    // Multilinear.init(table_evals) :124 -> EmptyEvaluations / LengthNotPowerOfTwo
    if (n_table == 0) return ZB_ERR_EMPTY_EVALUATIONS;
    if (n_table & (n_table - 1)) return ZB_ERR_LENGTH_NOT_POW2;
    zb_mle table_poly = 0;
    int32_t rc = zb_xxh3_rows(ctx, table_rows, n_table, arity, n_table, &table_poly); // :119-122
    if (rc) return rc;
    return lasso_run(ctx, table_poly, query_rows, n_queries, arity, round_polys, final_point, final_eval, num_vars, qc, tc);
}

int32_t zh_lasso_prove_with_mapping(zb_ctx *ctx, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows,
                                    uint64_t n_queries, const uint64_t *mapping, uint64_t n_mapping, uint32_t arity,
                                    uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                                    uint8_t qc[32], uint8_t tc[32]) {
    if (n_queries != n_mapping) return ZB_ERR_MAPPING_LEN_MISMATCH; // :186-188
    // :191-201 — an O(#queries * arity) host comparison over data that is already on the host
    for (uint64_t j = 0; j < n_queries; j++) {
        if (mapping[j] >= n_table) return ZB_ERR_INVALID_MAPPING;
        if (memcmp(query_rows + j * arity, table_rows + mapping[j] * arity, arity * sizeof(uint64_t)) != 0)
            return ZB_ERR_QUERY_TABLE_MISMATCH; // entriesMatch :255-268
    }
    return zh_lasso_prove(ctx, table_rows, n_table, query_rows, n_queries, arity, round_polys, final_point, final_eval, num_vars,
                          qc, tc);
}

int32_t zh_lasso_prove_builtin(zb_ctx *ctx, int32_t op, uint32_t bits, const uint64_t *query_rows, uint64_t n_queries,
                               uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                               uint8_t qc[32], uint8_t tc[32]) {
    if (n_queries == 0) return ZB_ERR_NO_QUERIES;
    zb_mle table_poly = 0;
    int32_t rc = zb_table_mle(ctx, op, bits, &table_poly);
    if (rc) return rc;
    return lasso_run(ctx, table_poly, query_rows, n_queries, 3, round_polys, final_point, final_eval, num_vars, qc, tc);
}

int32_t zh_lasso_prove_builtin_batch(zb_ctx *ctx, uint32_t n_jobs, const int32_t *ops, const uint32_t *bits,
                                     const uint64_t *const *query_rows, const uint64_t *n_queries, uint64_t *const *round_polys,
                                     uint64_t *const *final_points, uint64_t *final_evals, uint32_t *num_vars,
                                     uint8_t *query_commitments, uint8_t *table_commitments, int32_t *statuses) {
    if (n_jobs == 0) return ZB_OK;
    if (!ops || !bits || !query_rows || !n_queries || !round_polys || !final_points || !final_evals || !num_vars ||
        !query_commitments || !table_commitments || !statuses)
        return ZB_ERR_BAD_ARGUMENT;
    const unsigned hw = std::thread::hardware_concurrency();
    const uint32_t wave = hw > 2 ? hw - 1 : 1; // one sponge thread per proof in flight, this thread drives the GPU
    int32_t first = ZB_OK;
    for (uint32_t base = 0; base < n_jobs; base += wave) {
        const uint32_t cnt = n_jobs - base < wave ? n_jobs - base : wave;
        std::vector<uint64_t> padded(cnt), off(cnt);
        std::vector<char> piped(cnt);
        uint64_t total = 0;
        for (uint32_t k = 0; k < cnt; k++) {
            const uint64_t nq = n_queries[base + k];
            padded[k] = ceil_pow2(nq);
            piped[k] = nq > 0 && lasso_pipeline(padded[k]);
            off[k] = total;
            if (piped[k]) total += padded[k];
        }
        uint32_t *mirror = nullptr;
        int32_t mrc = total ? zb_host_mirror(ctx, total * sizeof(uint32_t), (void **)&mirror) : ZB_OK;
        std::vector<std::unique_ptr<SpongeFeed>> feeds(cnt);
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t j = base + k;
            uint8_t *qc = query_commitments + 32 * (size_t)j, *tc = table_commitments + 32 * (size_t)j;
            if (!piped[k]) {
                statuses[j] = zh_lasso_prove_builtin(ctx, ops[j], bits[j], query_rows[j], n_queries[j], round_polys[j],
                                                     final_points[j], final_evals + j, num_vars + j, qc, tc);
                continue;
            }
            statuses[j] = mrc;
            if (mrc) continue;
            zb_mle table_poly = 0;
            statuses[j] = zb_table_mle(ctx, ops[j], bits[j], &table_poly);
            if (statuses[j]) continue;
            feeds[k].reset(new SpongeFeed);
            const bool threaded = feeds[k]->start(mirror + off[k], padded[k]);
            statuses[j] = lasso_piped_body(ctx, table_poly, query_rows[j], n_queries[j], 3, padded[k], mirror + off[k], *feeds[k],
                                           round_polys[j], final_points[j], final_evals + j, num_vars + j, tc);
            zb_mle_free(ctx, table_poly);
            if (!threaded && statuses[j] == ZB_OK) feeds[k]->run();
            if (statuses[j]) feeds[k]->finish(false, nullptr); // stop this sponge now; the others keep running
        }
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t j = base + k;
            if (feeds[k] && statuses[j] == ZB_OK) feeds[k]->finish(true, query_commitments + 32 * (size_t)j);
            if (statuses[j] && first == ZB_OK) first = statuses[j];
        }
    }
    return first;
}

} // extern "C"
