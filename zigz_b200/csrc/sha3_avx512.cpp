// Keccak-f[1600] absorb loop with AVX-512 for the long sequential host sponges of the path
// (LassoProver.commitToPolynomial, /root/reference/src/lookups/lasso_prover.zig:242-252: one SHA3-256 over 8n bytes,
//  n/17 DEPENDENT permutations — the largest serial cost of a Lasso proof). One state, five zmm registers:
// plane y in register y, lane x in 64-bit slot x. Per round:
//   theta : C = xor5(planes) (2 vpternlogq); D folded into each plane with one vpternlogq
//   rho   : one vprolvq per plane
//   pi    : one vpermq per plane moves every lane to the slot of its destination PLANE; chi then works across
//           registers (5 vpternlogq, imm 0xD2) and leaves the state transposed (register = x, slot = y)
//   iota, then a 5x5 transpose back to planes: 4 unpacks, 2 vpermt2q that park e4's lanes in the spare slots of the
//   (e2, e3) unpacks, 5 vpermt2q, 1 blend. 18 cross-lane shuffles per round, all on one port: measured 33 cycles per
//   round on the pool's hosts against 37 for a transpose with 14 shuffles (the shuffle chain is the critical path).
// Compiled with -mavx512f; selected at run time (sha3_host.cpp) only when the CPU reports AVX-512F.
#include <cstddef>
#include <cstdint>
#include <immintrin.h>

namespace zigz {

namespace {
alignas(64) const uint64_t RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
} // namespace

namespace {
struct LoadU64 {
    static __m512i load(__mmask8 m, const uint64_t *p) { return _mm512_maskz_loadu_epi64(m, p); }
};
struct LoadU32 { // canonical u32 field elements, absorbed as their 8-byte little-endian encodings
    static __m512i load(__mmask8 m, const uint32_t *p) { return _mm512_cvtepu32_epi64(_mm256_maskz_loadu_epi32(m, p)); }
};
} // namespace

template <typename W, typename L>
static inline void absorb_impl(uint64_t state[25], const W *words, size_t nblocks, size_t lanes_per_block) {
    // lanes_per_block = 17 for SHA3-256 (rate 136)
    const __m512i rho0 = _mm512_setr_epi64(0, 1, 62, 28, 27, 0, 0, 0);
    const __m512i rho1 = _mm512_setr_epi64(36, 44, 6, 55, 20, 0, 0, 0);
    const __m512i rho2 = _mm512_setr_epi64(3, 10, 43, 25, 39, 0, 0, 0);
    const __m512i rho3 = _mm512_setr_epi64(41, 45, 15, 21, 8, 0, 0, 0);
    const __m512i rho4 = _mm512_setr_epi64(18, 2, 61, 56, 14, 0, 0, 0);
    const __m512i prev = _mm512_setr_epi64(4, 0, 1, 2, 3, 5, 6, 7); // C[x-1]
    const __m512i next = _mm512_setr_epi64(1, 2, 3, 4, 0, 5, 6, 7); // C[x+1]
    // pi: Q_i[j] = P_i[(3j + i) mod 5]
    const __m512i pi0 = _mm512_setr_epi64(0, 3, 1, 4, 2, 5, 6, 7);
    const __m512i pi1 = _mm512_setr_epi64(1, 4, 2, 0, 3, 5, 6, 7);
    const __m512i pi2 = _mm512_setr_epi64(2, 0, 3, 1, 4, 5, 6, 7);
    const __m512i pi3 = _mm512_setr_epi64(3, 1, 4, 2, 0, 5, 6, 7);
    const __m512i pi4 = _mm512_setr_epi64(4, 2, 0, 3, 1, 5, 6, 7);
    // transpose helpers: e4's lanes ride in the two spare slots of the (e2, e3) unpacks
    const __m512i injL = _mm512_setr_epi64(0, 1, 2, 3, 4, 5, 8 + 0, 8 + 2);
    const __m512i injH = _mm512_setr_epi64(0, 1, 2, 3, 4, 5, 8 + 1, 8 + 3);
    const __m512i tA = _mm512_setr_epi64(0, 1, 8, 9, 14, 0, 0, 0);
    const __m512i tB = _mm512_setr_epi64(2, 3, 10, 11, 15, 0, 0, 0);
    const __m512i tC = _mm512_setr_epi64(4, 5, 12, 13, 0, 0, 0, 0);

    __m512i p0 = _mm512_maskz_loadu_epi64(0x1F, state + 0);
    __m512i p1 = _mm512_maskz_loadu_epi64(0x1F, state + 5);
    __m512i p2 = _mm512_maskz_loadu_epi64(0x1F, state + 10);
    __m512i p3 = _mm512_maskz_loadu_epi64(0x1F, state + 15);
    __m512i p4 = _mm512_maskz_loadu_epi64(0x1F, state + 20);

    for (size_t blk = 0; blk < nblocks; blk++, words += lanes_per_block) {
        // absorb (lanes_per_block == 17: planes 0..2 fully, plane 3 lanes 15, 16)
        p0 = _mm512_xor_si512(p0, L::load(0x1F, words + 0));
        p1 = _mm512_xor_si512(p1, L::load(0x1F, words + 5));
        p2 = _mm512_xor_si512(p2, L::load(0x1F, words + 10));
        p3 = _mm512_xor_si512(p3, L::load(0x03, words + 15));
        for (int r = 0; r < 24; r++) {
            // theta
            __m512i c = _mm512_ternarylogic_epi64(_mm512_ternarylogic_epi64(p0, p1, p2, 0x96), p3, p4, 0x96);
            const __m512i cm = _mm512_permutexvar_epi64(prev, c);
            const __m512i cp = _mm512_rol_epi64(_mm512_permutexvar_epi64(next, c), 1);
            p0 = _mm512_ternarylogic_epi64(p0, cm, cp, 0x96);
            p1 = _mm512_ternarylogic_epi64(p1, cm, cp, 0x96);
            p2 = _mm512_ternarylogic_epi64(p2, cm, cp, 0x96);
            p3 = _mm512_ternarylogic_epi64(p3, cm, cp, 0x96);
            p4 = _mm512_ternarylogic_epi64(p4, cm, cp, 0x96);
            // rho + pi (to the slot of the destination plane)
            const __m512i q0 = _mm512_permutexvar_epi64(pi0, _mm512_rolv_epi64(p0, rho0));
            const __m512i q1 = _mm512_permutexvar_epi64(pi1, _mm512_rolv_epi64(p1, rho1));
            const __m512i q2 = _mm512_permutexvar_epi64(pi2, _mm512_rolv_epi64(p2, rho2));
            const __m512i q3 = _mm512_permutexvar_epi64(pi3, _mm512_rolv_epi64(p3, rho3));
            const __m512i q4 = _mm512_permutexvar_epi64(pi4, _mm512_rolv_epi64(p4, rho4));
            // chi across registers: e_i[j] = new lane (x = i, y = j)
            __m512i e0 = _mm512_ternarylogic_epi64(q0, q1, q2, 0xD2);
            const __m512i e1 = _mm512_ternarylogic_epi64(q1, q2, q3, 0xD2);
            const __m512i e2 = _mm512_ternarylogic_epi64(q2, q3, q4, 0xD2);
            const __m512i e3 = _mm512_ternarylogic_epi64(q3, q4, q0, 0xD2);
            const __m512i e4 = _mm512_ternarylogic_epi64(q4, q0, q1, 0xD2);
            // iota: lane (0, 0) = e0 slot 0
            e0 = _mm512_xor_si512(e0, _mm512_maskz_set1_epi64(0x01, (long long)RC[r]));
            // transpose back to planes: p_j[i] = e_i[j] (slots 5..7 of the planes carry don't-care values)
            const __m512i lo01 = _mm512_unpacklo_epi64(e0, e1), hi01 = _mm512_unpackhi_epi64(e0, e1);
            const __m512i lo23 = _mm512_permutex2var_epi64(_mm512_unpacklo_epi64(e2, e3), injL, e4);
            const __m512i hi23 = _mm512_permutex2var_epi64(_mm512_unpackhi_epi64(e2, e3), injH, e4);
            p0 = _mm512_permutex2var_epi64(lo01, tA, lo23);
            p1 = _mm512_permutex2var_epi64(hi01, tA, hi23);
            p2 = _mm512_permutex2var_epi64(lo01, tB, lo23);
            p3 = _mm512_permutex2var_epi64(hi01, tB, hi23);
            p4 = _mm512_mask_blend_epi64(0x10, _mm512_permutex2var_epi64(lo01, tC, lo23), e4);
        }
    }
    _mm512_mask_storeu_epi64(state + 0, 0x1F, p0);
    _mm512_mask_storeu_epi64(state + 5, 0x1F, p1);
    _mm512_mask_storeu_epi64(state + 10, 0x1F, p2);
    _mm512_mask_storeu_epi64(state + 15, 0x1F, p3);
    _mm512_mask_storeu_epi64(state + 20, 0x1F, p4);
}

void keccak_absorb_avx512(uint64_t state[25], const uint64_t *words, size_t nblocks, size_t lanes_per_block) {
    absorb_impl<uint64_t, LoadU64>(state, words, nblocks, lanes_per_block);
}

void keccak_absorb_u32_avx512(uint64_t state[25], const uint32_t *words, size_t nblocks) {
    absorb_impl<uint32_t, LoadU32>(state, words, nblocks, 17);
}

} // namespace zigz
