/* oracle/zo_hash.c — TEST INFRASTRUCTURE ONLY. See zo_hash.h for provenance. */
#include "zo_hash.h"
#include <string.h>

/* ------------------------------------------------------------------ */
/* Keccak-f[1600] / SHA3-256 (FIPS 202)                                */
/* ------------------------------------------------------------------ */
static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14,
                                   27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4,
                                   15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static inline uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

void zo_keccak_f1600(uint64_t s[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5], t;
        for (int i = 0; i < 5; i++) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
        for (int i = 0; i < 5; i++) {
            t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
        }
        t = s[1];
        for (int i = 0; i < 24; i++) {
            int j = KECCAK_PIL[i];
            uint64_t b = s[j];
            s[j] = rotl64(t, KECCAK_ROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = s[j + i];
            for (int i = 0; i < 5; i++) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        s[0] ^= KECCAK_RC[round];
    }
}

static inline uint64_t load_le64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}
static inline uint32_t load_le32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

static void sha3_absorb_block(uint64_t s[25], const uint8_t *blk) {
    for (int i = 0; i < 17; i++) s[i] ^= load_le64(blk + 8 * i);
    zo_keccak_f1600(s);
}

void zo_sha3_256_init(zo_sha3_256 *c) { memset(c, 0, sizeof(*c)); }

void zo_sha3_256_update(zo_sha3_256 *c, const void *data, size_t len) {
    const uint8_t *p = (const uint8_t *)data;
    if (c->buf_len) {
        size_t take = 136 - c->buf_len;
        if (take > len) take = len;
        memcpy(c->buf + c->buf_len, p, take);
        c->buf_len += (uint32_t)take;
        p += take;
        len -= take;
        if (c->buf_len == 136) {
            sha3_absorb_block(c->s, c->buf);
            c->buf_len = 0;
        }
    }
    while (len >= 136) {
        sha3_absorb_block(c->s, p);
        p += 136;
        len -= 136;
    }
    if (len) {
        memcpy(c->buf, p, len);
        c->buf_len = (uint32_t)len;
    }
}

void zo_sha3_256_peek(const zo_sha3_256 *c, uint8_t out[32]) {
    uint64_t s[25];
    uint8_t blk[136];
    memcpy(s, c->s, sizeof(s));
    memset(blk, 0, sizeof(blk));
    memcpy(blk, c->buf, c->buf_len);
    blk[c->buf_len] ^= 0x06; /* SHA-3 domain bits + first pad bit */
    blk[135] ^= 0x80;
    sha3_absorb_block(s, blk);
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(s[i] >> (8 * b));
}

void zo_sha3_256_oneshot(const void *data, size_t len, uint8_t out[32]) {
    zo_sha3_256 c;
    zo_sha3_256_init(&c);
    zo_sha3_256_update(&c, data, len);
    zo_sha3_256_peek(&c, out);
}

/* ------------------------------------------------------------------ */
/* XXH3-64, inputs of 0..16 bytes (XXH3 spec, default 192-byte secret)  */
/* ------------------------------------------------------------------ */
static const uint8_t XXH3_SECRET[72] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb};

#define XXH_PRIME64_2 0xC2B2AE3D27D4EB4FULL
#define XXH_PRIME64_3 0x165667B19E3779F9ULL

static inline uint64_t xxh64_avalanche(uint64_t h) {
    h ^= h >> 33;
    h *= XXH_PRIME64_2;
    h ^= h >> 29;
    h *= XXH_PRIME64_3;
    h ^= h >> 32;
    return h;
}
static inline uint64_t xxh3_avalanche(uint64_t h) {
    h ^= h >> 37;
    h *= 0x165667919E3779F9ULL;
    h ^= h >> 32;
    return h;
}
static inline uint64_t bswap64(uint64_t x) { return __builtin_bswap64(x); }
static inline uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }

uint64_t zo_xxh3_64_small(const void *data, size_t len, uint64_t seed) {
    const uint8_t *in = (const uint8_t *)data;
    const uint8_t *sec = XXH3_SECRET;
    if (len == 0) return xxh64_avalanche(seed ^ (load_le64(sec + 56) ^ load_le64(sec + 64)));
    if (len <= 3) {
        uint8_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t combined = ((uint32_t)c1 << 16) | ((uint32_t)c2 << 24) | (uint32_t)c3 | ((uint32_t)len << 8);
        uint64_t bitflip = (uint64_t)(load_le32(sec) ^ load_le32(sec + 4)) + seed;
        return xxh64_avalanche((uint64_t)combined ^ bitflip);
    }
    if (len <= 8) {
        seed ^= (uint64_t)bswap32((uint32_t)seed) << 32;
        uint32_t in1 = load_le32(in), in2 = load_le32(in + len - 4);
        uint64_t bitflip = (load_le64(sec + 8) ^ load_le64(sec + 16)) - seed;
        uint64_t in64 = (uint64_t)in2 + ((uint64_t)in1 << 32);
        uint64_t h = in64 ^ bitflip;
        /* XXH3_rrmxmx */
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= 0x9FB21C651E98DF25ULL;
        h ^= (h >> 35) + (uint64_t)len;
        h *= 0x9FB21C651E98DF25ULL;
        return h ^ (h >> 28);
    }
    /* 9..16 */
    {
        uint64_t bitflip1 = (load_le64(sec + 24) ^ load_le64(sec + 32)) + seed;
        uint64_t bitflip2 = (load_le64(sec + 40) ^ load_le64(sec + 48)) - seed;
        uint64_t lo = load_le64(in) ^ bitflip1;
        uint64_t hi = load_le64(in + len - 8) ^ bitflip2;
        unsigned __int128 m = (unsigned __int128)lo * hi;
        uint64_t acc = (uint64_t)len + bswap64(lo) + hi + ((uint64_t)m ^ (uint64_t)(m >> 64));
        return xxh3_avalanche(acc);
    }
}

/* ------------------------------------------------------------------ */
/* SHA-256 (FIPS 180-4)                                                */
/* ------------------------------------------------------------------ */
static const uint32_t SHA256_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_block(uint32_t h[8], const uint8_t *p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA256_K[i] + w[i];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void zo_sha256_oneshot(const void *data, size_t len, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const uint8_t *p = (const uint8_t *)data;
    size_t n = len;
    while (n >= 64) {
        sha256_block(h, p);
        p += 64;
        n -= 64;
    }
    uint8_t blk[128];
    memset(blk, 0, sizeof(blk));
    memcpy(blk, p, n);
    blk[n] = 0x80;
    size_t tot = (n + 9 <= 64) ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) blk[tot - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(h, blk);
    if (tot == 128) sha256_block(h, blk + 64);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8);
        out[4 * i + 3] = (uint8_t)h[i];
    }
}
