/* oracle/zo_hash.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C restatements of the three standard hashes the reference's proving
 * path uses through Zig's standard library:
 *   - SHA3-256  (std.crypto.hash.sha3.Sha3_256)  /root/reference/src/core/hash.zig:135-147,187-195,255-316
 *   - XXH3-64   (std.hash.XxHash3.hash(seed, bytes)) /root/reference/src/lookups/lasso_prover.zig:208-239
 *   - SHA-256   (std.crypto.hash.sha2.Sha256)     /root/reference/src/prover/prover.zig:99
 * The Zig std implementations are not part of /root/reference; the algorithms
 * restated here are the published ones (FIPS 202, XXH3 spec v0.8, FIPS 180-4)
 * and the tests pin them against hashlib / the xxhash package.
 */
#ifndef ZO_HASH_H
#define ZO_HASH_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t s[25];
    uint8_t buf[136]; /* rate of SHA3-256 */
    uint32_t buf_len;
} zo_sha3_256;

void zo_keccak_f1600(uint64_t s[25]);
void zo_sha3_256_init(zo_sha3_256 *c);
void zo_sha3_256_update(zo_sha3_256 *c, const void *data, size_t len);
/* finalises a COPY of the state: the streaming context stays usable, which is
 * what FiatShamirTranscript.challenge needs (hash.zig:304-306). */
void zo_sha3_256_peek(const zo_sha3_256 *c, uint8_t out[32]);
void zo_sha3_256_oneshot(const void *data, size_t len, uint8_t out[32]);

/* XXH3_64bits_withSeed for inputs of 0..16 bytes (the path only ever hashes 8). */
uint64_t zo_xxh3_64_small(const void *data, size_t len, uint64_t seed);

void zo_sha256_oneshot(const void *data, size_t len, uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif
