// Measured integer-pipe ceiling for the hashing kernels (SURVEY.md §8d asks for a MEASURED LOP3/SHF peak next to the
// Merkle numbers): register-only kernels issuing independent LOP3 (chi-like, non-linear so nothing folds away) and
// SHF.L.W chains, 16 chains per thread for ILP, enough CTAs to fill every SM. Reports 32-bit lane-operations per second.
#include "kernels.h"

namespace zk {

__device__ __forceinline__ uint32_t op_lop3(uint32_t a, uint32_t b, uint32_t c) { // a ^ (~b & c)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t op_shf(uint32_t lo, uint32_t hi) {
    uint32_t d;
    asm("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(d) : "r"(lo), "r"(hi));
    return d;
}

// MODE 0: LOP3 only, 1: SHF only, 2: Keccak mix (per 16 registers: 11 LOP3 + 5 SHF ~ 122 : 58)
template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t *out, int iters, uint32_t seed) {
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = seed * (i + 1) + threadIdx.x + blockIdx.x * 977u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0 || (MODE == 2 && (i % 16) < 11))
                    x[i] = op_lop3(x[i], x[(i + 1) & 15], x[(i + 2) & 15]);
                else
                    x[i] = op_shf(x[i], x[(i + 5) & 15]);
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    if (acc == 0x12345678u) out[0] = acc; // practically never: keeps the chains alive
}

// returns lane-operations per second for the three modes
void launch_int_peak(int mode, uint32_t *out, int iters, int ctas, cudaStream_t st) {
    if (mode == 0) k_int_peak<0><<<ctas, 256, 0, st>>>(out, iters, 12345u);
    else if (mode == 1) k_int_peak<1><<<ctas, 256, 0, st>>>(out, iters, 12345u);
    else k_int_peak<2><<<ctas, 256, 0, st>>>(out, iters, 12345u);
}

} // namespace zk
