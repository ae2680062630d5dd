// Keccak-f[1600] on 32-bit register pairs for sm_100a (no 64-bit integer ALU on the SM).
// SHA3-256 as used by the reference: std.crypto.hash.sha3.Sha3_256 in
// /root/reference/src/core/hash.zig:135-147 (leaf) and :187-195 (node).
//   theta : 5-input XOR columns as two LOP3 (lut 0x96); D is never materialised:
//           A ^ C[x-1] ^ rotl(C[x+1],1) is one LOP3 per half lane
//   rho/pi: one SHF.L.W funnel shift per half lane, compile-time amounts
//   chi   : a ^ (~b & c) = one LOP3 (lut 0xD2) per half lane
// => ~180 ALU-pipe instructions per round, 24 rounds.
#pragma once
#include <cstdint>

namespace keccak {

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t chi(uint32_t a, uint32_t b, uint32_t c) { // a ^ (~b & c)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

static __device__ __constant__ const uint32_t RC_LO[24] = {
    0x00000001u, 0x00008082u, 0x0000808au, 0x80008000u, 0x0000808bu, 0x80000001u, 0x80008081u, 0x00008009u,
    0x0000008au, 0x00000088u, 0x80008009u, 0x8000000au, 0x8000808bu, 0x0000008bu, 0x00008089u, 0x00008003u,
    0x00008002u, 0x00000080u, 0x0000800au, 0x8000000au, 0x80008081u, 0x00008080u, 0x80000001u, 0x80008008u};
static __device__ __constant__ const uint32_t RC_HI[24] = {
    0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u,
    0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u,
    0x80000000u, 0x80000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x00000000u, 0x80000000u};

// rho offsets indexed by lane x + 5y
__device__ constexpr int RHO[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                                    25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};

// 2^k multipliers for the FMA-pipe rotations. Filled once per context from the host (keccak_init_constants) so that
// ptxas cannot strength-reduce the multiplications back into ALU-pipe shifts.
static __device__ __constant__ uint32_t POW2[33];

// rotl64 by N (1..31 after the free half swap) on the FMA pipe: x * 2^N as a 96-bit product, the bits shifted out at
// the top re-enter at the bottom (disjoint bit ranges, so the final OR is an addition):
//   p0     = lo * 2^N                  IMAD.WIDE.U32      p0.lo = lo << N, p0.hi = lo >> (32-N)
//   new hi = lo32(hi * 2^N) + p0.hi    IMAD
//   new lo = hi32(hi * 2^N) + p0.lo    IMAD.HI.U32
// 3 FMA-pipe instructions instead of 2 ALU-pipe funnel shifts: the Keccak kernels are ALU-pipe bound with the FMA
// pipe idle (profiles/r01_ncu_full_merkle_2p24.txt), so moving a share of the rotations rebalances the two pipes.
template <int N>
__device__ __forceinline__ void rotl64_fma(uint32_t lo, uint32_t hi, uint32_t &olo, uint32_t &ohi) {
    static_assert(N >= 1 && N <= 63 && N != 32, "rotation amount");
    constexpr int K = N & 31;
    const uint32_t a = N < 32 ? lo : hi, b = N < 32 ? hi : lo; // N >= 32: swap halves first
    const uint32_t m = POW2[K];
    uint32_t p0l, p0h;
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(p0l), "=r"(p0h) : "r"(a), "r"(m));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ohi) : "r"(b), "r"(m), "r"(p0h));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(olo) : "r"(b), "r"(m), "r"(p0l));
}

template <int N>
__device__ __forceinline__ void rotl64(uint32_t lo, uint32_t hi, uint32_t &olo, uint32_t &ohi) {
    if constexpr (N == 0) {
        olo = lo;
        ohi = hi;
    } else if constexpr (N < 32) {
        ohi = __funnelshift_l(lo, hi, N);
        olo = __funnelshift_l(hi, lo, N);
    } else if constexpr (N == 32) {
        olo = hi;
        ohi = lo;
    } else {
        ohi = __funnelshift_l(hi, lo, N - 32);
        olo = __funnelshift_l(lo, hi, N - 32);
    }
}

template <int I, uint32_t FMAMASK>
__device__ __forceinline__ void rho_pi_lane(const uint32_t (&al)[25], const uint32_t (&ah)[25], const uint32_t (&dl)[5],
                                            const uint32_t (&dh)[5], const uint32_t (&el)[5], const uint32_t (&eh)[5],
                                            uint32_t (&bl)[25], uint32_t (&bh)[25]) {
    constexpr int x = I % 5, y = I / 5;
    constexpr int dst = y + 5 * ((2 * x + 3 * y) % 5);
    // theta folded in: t = A ^ C[x-1] ^ rotl(C[x+1], 1)
    uint32_t tl = xor3(al[I], dl[x], el[x]);
    uint32_t th = xor3(ah[I], dh[x], eh[x]);
    if constexpr (((FMAMASK >> I) & 1u) != 0 && RHO[I] != 0)
        rotl64_fma<RHO[I]>(tl, th, bl[dst], bh[dst]);
    else
        rotl64<RHO[I]>(tl, th, bl[dst], bh[dst]);
    if constexpr (I + 1 < 25) rho_pi_lane<I + 1, FMAMASK>(al, ah, dl, dh, el, eh, bl, bh);
}

// FMAMASK: bit i set -> the rho rotation of lane i runs on the FMA pipe (bits 25..29: the five theta rotl-by-1)
template <uint32_t FMAMASK>
__device__ __forceinline__ void round(uint32_t (&al)[25], uint32_t (&ah)[25], uint32_t rcl, uint32_t rch) {
    uint32_t cl[5], ch[5];
#pragma unroll
    for (int x = 0; x < 5; x++) {
        cl[x] = xor3(xor3(al[x], al[x + 5], al[x + 10]), al[x + 15], al[x + 20]);
        ch[x] = xor3(xor3(ah[x], ah[x + 5], ah[x + 10]), ah[x + 15], ah[x + 20]);
    }
    // dl/dh = C[x-1]; el/eh = rotl(C[x+1], 1)
    uint32_t dl[5], dh[5], el[5], eh[5];
#pragma unroll
    for (int x = 0; x < 5; x++) {
        dl[x] = cl[(x + 4) % 5];
        dh[x] = ch[(x + 4) % 5];
        if ((FMAMASK >> (25 + x)) & 1u)
            rotl64_fma<1>(cl[(x + 1) % 5], ch[(x + 1) % 5], el[x], eh[x]);
        else
            rotl64<1>(cl[(x + 1) % 5], ch[(x + 1) % 5], el[x], eh[x]);
    }
    uint32_t bl[25], bh[25];
    rho_pi_lane<0, FMAMASK>(al, ah, dl, dh, el, eh, bl, bh);
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
#pragma unroll
        for (int x = 0; x < 5; x++) {
            al[y + x] = chi(bl[y + x], bl[y + (x + 1) % 5], bl[y + (x + 2) % 5]);
            ah[y + x] = chi(bh[y + x], bh[y + (x + 1) % 5], bh[y + (x + 2) % 5]);
        }
    }
    al[0] ^= rcl;
    ah[0] ^= rch;
}

// UNROLL = rounds per loop iteration (24 = fully unrolled, constants folded)
template <int UNROLL, uint32_t FMAMASK = 0>
__device__ __forceinline__ void f1600(uint32_t (&al)[25], uint32_t (&ah)[25]) {
    if constexpr (UNROLL > 100) {
        // PEELED form, UNROLL = 100 + rounds per iteration of the middle loop (2 or 11): round 0 and round 23 stand outside the
        // loop as straight-line code, so the compiler folds the constant lanes of the padded input block into round 0 (17 of
        // the 25 lanes of a node block, 24 of a leaf block) and strips round 23 down to the four output lanes (chi of row 0
        // only: 5 of the 25 rho/pi lanes) — neither is possible inside a rolled loop.
        constexpr int UR = UNROLL - 100;
        static_assert(22 % UR == 0, "the middle 22 rounds must split evenly");
        round<FMAMASK>(al, ah, RC_LO[0], RC_HI[0]);
#pragma unroll 1
        for (int r = 1; r < 23; r += UR) {
#pragma unroll
            for (int k = 0; k < UR; k++) round<FMAMASK>(al, ah, RC_LO[r + k], RC_HI[r + k]);
        }
        round<FMAMASK>(al, ah, RC_LO[23], RC_HI[23]);
    } else if constexpr (UNROLL >= 24) {
#pragma unroll
        for (int r = 0; r < 24; r++) round<FMAMASK>(al, ah, RC_LO[r], RC_HI[r]);
    } else {
#pragma unroll 1
        for (int r = 0; r < 24; r += UNROLL) {
#pragma unroll
            for (int k = 0; k < UNROLL; k++) round<FMAMASK>(al, ah, RC_LO[r + k], RC_HI[r + k]);
        }
    }
}

// SHA3-256 of the 8-byte little-endian encoding of a field element (leaf hash, merkle_tree.zig:298-300)
template <int UNROLL, uint32_t FMAMASK = 0>
__device__ __forceinline__ void sha3_leaf(uint32_t value, uint32_t (&out)[8]) {
    uint32_t al[25], ah[25];
#pragma unroll
    for (int i = 0; i < 25; i++) {
        al[i] = 0;
        ah[i] = 0;
    }
    al[0] = value;       // le64(value), high word is zero: value < 2^31
    al[1] = 0x06u;       // SHA-3 domain separation + first pad bit right after the 8 message bytes
    ah[16] = 0x80000000u; // last pad bit: byte 135 of the 136-byte rate
    f1600<UNROLL, FMAMASK>(al, ah);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        out[2 * i] = al[i];
        out[2 * i + 1] = ah[i];
    }
}

// SHA3-256(left || right) of two 32-byte digests (node hash, merkle_tree.zig:390-392)
template <int UNROLL, uint32_t FMAMASK = 0>
__device__ __forceinline__ void sha3_node(const uint32_t (&in)[16], uint32_t (&out)[8]) {
    uint32_t al[25], ah[25];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        al[i] = in[2 * i];
        ah[i] = in[2 * i + 1];
    }
#pragma unroll
    for (int i = 8; i < 25; i++) {
        al[i] = 0;
        ah[i] = 0;
    }
    al[8] = 0x06u;
    ah[16] = 0x80000000u;
    f1600<UNROLL, FMAMASK>(al, ah);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        out[2 * i] = al[i];
        out[2 * i + 1] = ah[i];
    }
}

} // namespace keccak
