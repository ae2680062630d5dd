// Keccak-f[1600] for the host-resident sponges (see sha3_host.hpp).
#include "sha3_host.hpp"

#include <cstdlib>

namespace zigz {

namespace {
constexpr uint64_t RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};

// rho rotation of lane (x, y), index x + 5y
constexpr int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

inline uint64_t rol(uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }

// one round: in -> out (out lanes are written in chi order so `in` can be reused as the next round's output)
inline __attribute__((always_inline)) void round_fn(const uint64_t *__restrict in, uint64_t *__restrict out, uint64_t rc) {
    uint64_t c[5], d[5], b[25];
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) c[x] = in[x] ^ in[x + 5] ^ in[x + 10] ^ in[x + 15] ^ in[x + 20];
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
#pragma GCC unroll 5
    for (int y = 0; y < 5; y++) {
#pragma GCC unroll 5
        for (int x = 0; x < 5; x++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(in[x + 5 * y] ^ d[x], RHO[x + 5 * y]);
    }
#pragma GCC unroll 5
    for (int y = 0; y < 5; y++) {
#pragma GCC unroll 5
        for (int x = 0; x < 5; x++) out[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    }
    out[0] ^= rc;
}
} // namespace

void keccak_absorb_avx512(uint64_t state[25], const uint64_t *words, size_t nblocks, size_t lanes_per_block);
void keccak_absorb_u32_avx512(uint64_t state[25], const uint32_t *words, size_t nblocks);

bool Sha3_256::have_avx512() {
    static const bool ok = [] {
        const char *e = getenv("ZB_HOST_SHA3"); // "scalar" forces the portable path
        if (e && !strcmp(e, "scalar")) return false;
        __builtin_cpu_init();
        return (bool)__builtin_cpu_supports("avx512f");
    }();
    return ok;
}

void Sha3_256::absorb_blocks_avx512(uint64_t s[25], const uint64_t *words, size_t nblocks) {
    keccak_absorb_avx512(s, words, nblocks, RATE / 8);
}

void Sha3_256::absorb_blocks_u32_avx512(uint64_t s[25], const uint32_t *words, size_t nblocks) {
    keccak_absorb_u32_avx512(s, words, nblocks);
}

void Sha3_256::permute(uint64_t s[25]) {
    uint64_t t[25];
    for (int r = 0; r < 24; r += 2) {
        round_fn(s, t, RC[r]);
        round_fn(t, s, RC[r + 1]);
    }
}

} // namespace zigz
