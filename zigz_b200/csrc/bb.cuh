// BabyBear arithmetic in registers (p = 2^31 - 2^27 + 1), canonical u32 in / canonical u32 out.
// Replaces Field(u64, 2013265921).{add,sub,mul} — /root/reference/src/core/field.zig:73-147.
// The reference multiplies with a u128 `%`; any exact modular product gives the same canonical value,
// so the device uses Shoup (fixed multiplier) and Montgomery (general) reductions on the 32-bit IMAD pipe.
#pragma once
#include <cstdint>

namespace bb {

constexpr uint32_t P = 2013265921u;        // 0x78000001
constexpr uint32_t P_NEG_INV = 2013265919u; // -P^{-1} mod 2^32  (P * 0x88000001 == 1 mod 2^32)
constexpr uint32_t R_MOD_P = 268435454u;   // 2^32 mod P
constexpr uint32_t R2_MOD_P = 1172168163u; // 2^64 mod P
constexpr uint32_t R3_MOD_P = 317946875u; // 2^96 mod P
constexpr uint32_t R4_MOD_P = 663890614u; // 2^128 mod P

__host__ __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return s >= P ? s - P : s;
}
__host__ __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    return a < b ? d + P : d;
}
__host__ __device__ __forceinline__ uint32_t reduce64(uint64_t x) { return (uint32_t)(x % P); }

// Shoup precomputation for a fixed multiplier w < P: w' = floor(w * 2^32 / P)
__host__ __device__ __forceinline__ uint32_t shoup_pre(uint32_t w) { return (uint32_t)(((uint64_t)w << 32) / P); }

// x * w mod P for any x < 2^32, w < P with w' = shoup_pre(w); 1 mulhi + 2 mullo
__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint32_t w, uint32_t wp) {
    uint32_t q = __umulhi(x, wp);
    uint32_t t = x * w - q * P; // in [0, 2P) and 2P < 2^32
    return t >= P ? t - P : t;
}

// Montgomery product a*b*2^-32 mod P, canonical output, for a*b < P*2^32
__device__ __forceinline__ uint32_t mont_mul(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a * b;
    uint32_t m = (uint32_t)t * P_NEG_INV;
    uint32_t u = (uint32_t)((t + (uint64_t)m * P) >> 32); // [0, 2P)
    return u >= P ? u - P : u;
}

// Lazy forms (fewer instructions in the HBM-bound kernels, which otherwise brush the issue limit under the power cap):
//   sub_lazy  : a - b + P in [1, 2P) for canonical a, b (2P < 2^32)
//   mont_mul_lazy : a*b*2^-32 mod P in [0, 2P), valid for a < 2^32 (lazy allowed) and b < P (canonical):
//                   t = a*b < 2^32 P so hi(t) < P
__host__ __device__ __forceinline__ uint32_t sub_lazy(uint32_t a, uint32_t b) { return a - b + P; }
__device__ __forceinline__ uint32_t mont_mul_lazy(uint32_t a, uint32_t b) {
    // subtractive form: with m = lo(t) * P^{-1}, t - m P is an exact multiple of 2^32 and (t - m P) / 2^32 = hi(t) - hi(m P),
    // no carry to propagate. hi(t) < P and hi(m P) < P, so adding P lands in [1, 2P).
    const uint64_t t = (uint64_t)a * b;
    const uint32_t m = (uint32_t)t * 0x88000001u; // P^{-1} mod 2^32
    return (uint32_t)(t >> 32) + P - __umulhi(m, P);
}

// t * R^E mod P for ANY 64-bit t (E = 0, 1, 2), canonical, without a 64-bit division: t = hi 2^32 + lo, and a Montgomery product
// with the constant R^(E+2) resp. R^(E+1) turns each half into its share (mont_mul divides by R once). This is how the raw
// u64 accumulators of the round kernels become canonical field elements; ~12 instructions where `t % P` (software division)
// followed by a modular product costs a few hundred — it runs on ONE thread per kernel (the publishing one), 16 times for a
// round grid, on the critical path of every host round trip.
template <int E>
__device__ __forceinline__ uint32_t reduce64_scaled(unsigned long long t) {
    constexpr uint32_t CH = E == 0 ? R2_MOD_P : (E == 1 ? R3_MOD_P : R4_MOD_P);
    constexpr uint32_t CL = E == 0 ? R_MOD_P : (E == 1 ? R2_MOD_P : R3_MOD_P);
    return add(mont_mul((uint32_t)(t >> 32), CH), mont_mul((uint32_t)t, CL));
}

// plain a*b mod P via two Montgomery steps is wasteful; for a one-off product use this
__host__ __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }

// linear interpolation lo + r*(hi - lo): the fold of multilinear.zig:166-173 with one modmul
__device__ __forceinline__ uint32_t lerp(uint32_t lo, uint32_t hi, uint32_t r, uint32_t rp) {
    return add(lo, mul_shoup(sub_lazy(hi, lo), r, rp)); // Shoup accepts any x < 2^32: no reduction of hi - lo needed
}

// Two variables bound in one step: w0*a0 + w1*a1 + w2*a2 + w3*a3 with the four bilinear weights
// ((1-r1)(1-r2), (1-r1)r2, r1(1-r2), r1 r2) given in Montgomery form (bilinear_weights). The four 62-bit products are
// summed in 64 bits (4 P^2 < 2^64) and reduced once: 11 instructions against 27 for three nested lerps, same canonical
// value (exact field arithmetic, field.zig:112-147).
constexpr uint32_t P_INV = 0x88000001u; // P^{-1} mod 2^32
__device__ __forceinline__ uint32_t dot4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t w0, uint32_t w1, uint32_t w2,
                                         uint32_t w3) {
    const uint64_t t = (uint64_t)a0 * w0 + (uint64_t)a1 * w1 + (uint64_t)a2 * w2 + (uint64_t)a3 * w3;
    uint32_t hi = (uint32_t)(t >> 32); // < 1.875 P
    hi = min(hi, hi - P);              // < P
    const uint32_t m = (uint32_t)t * P_INV;
    const uint32_t u = hi - __umulhi(m, P); // (t' - m P) / 2^32, exact; in (-P, P)
    return min(u, u + P);
}
// One variable bound as a dot product: w0*a0 + w1*a1 with ((1-r) R, r R); 2 P^2 < P 2^32, so hi(t) < P needs no correction.
__device__ __forceinline__ uint32_t dot2(uint32_t a0, uint32_t a1, uint32_t w0, uint32_t w1) {
    const uint64_t t = (uint64_t)a0 * w0 + (uint64_t)a1 * w1;
    const uint32_t m = (uint32_t)t * P_INV;
    const uint32_t u = (uint32_t)(t >> 32) - __umulhi(m, P);
    return min(u, u + P);
}
struct BilinearWeights {
    uint32_t w[4];
};
__host__ inline BilinearWeights bilinear_weights(uint32_t r1, uint32_t r2) {
    const uint32_t n1 = sub(1u, r1), n2 = sub(1u, r2);
    BilinearWeights b;
    b.w[0] = mul(mul(n1, n2), R_MOD_P);
    b.w[1] = mul(mul(n1, r2), R_MOD_P);
    b.w[2] = mul(mul(r1, n2), R_MOD_P);
    b.w[3] = mul(mul(r1, r2), R_MOD_P);
    return b;
}

} // namespace bb
