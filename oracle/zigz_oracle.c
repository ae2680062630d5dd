/* oracle/zigz_oracle.c — TEST INFRASTRUCTURE ONLY. See zigz_oracle.h.
 *
 * Each function follows the cited reference lines operation for operation
 * (same loop order, same modular formulae, same allocation pattern where it
 * matters for timing: per-round fresh array in partialEval, per-level arrays
 * and full recomputation in Merkle open), so that it can double as the
 * single-core "port" CPU baseline.
 */
#include "zigz_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Field(u64, p) — /root/reference/src/core/field.zig                   */
/* ------------------------------------------------------------------ */
uint64_t zo_f_init(uint64_t p, uint64_t v) { return v % p; } /* :36-38 */

uint64_t zo_f_add(uint64_t p, uint64_t a, uint64_t b) { /* :73-88 */
    uint64_t sum;
    if (__builtin_add_overflow(a, b, &sum)) return sum % p; /* reference quirk kept: wrapped value mod p */
    if (sum >= p) return sum - p;
    return sum;
}

uint64_t zo_f_sub(uint64_t p, uint64_t a, uint64_t b) { /* :91-98 */
    if (a >= b) return a - b;
    return p - (b - a);
}

uint64_t zo_f_neg(uint64_t p, uint64_t a) { return a == 0 ? 0 : p - a; } /* :101-106 */

uint64_t zo_f_mul(uint64_t p, uint64_t a, uint64_t b) { /* :112-147, the `bits <= 64` arm: u128 product % modulus */
    unsigned __int128 prod = (unsigned __int128)a * (unsigned __int128)b;
    return (uint64_t)(prod % (unsigned __int128)p);
}

uint64_t zo_f_pow(uint64_t p, uint64_t a, uint64_t e) { /* :204-225 */
    if (e == 0) return 1;
    if (e == 1) return a;
    uint64_t result = 1, base = a;
    while (e > 0) {
        if (e & 1) result = zo_f_mul(p, result, base);
        base = zo_f_mul(p, base, base);
        e >>= 1;
    }
    return result;
}

int zo_f_inv(uint64_t p, uint64_t a, uint64_t *out) { /* :157-191 extended Euclid in i128 */
    if (a == 0) return -1;
    __int128 t = 0, new_t = 1, r = (__int128)p, new_r = (__int128)a;
    while (new_r != 0) {
        __int128 q = r / new_r; /* both positive: divFloor == trunc */
        __int128 tmp = t;
        t = new_t;
        new_t = tmp - q * new_t;
        tmp = r;
        r = new_r;
        new_r = tmp - q * new_r;
    }
    if (r > 1) return -1;
    if (t < 0) t += (__int128)p;
    *out = (uint64_t)t;
    return 0;
}

/* ------------------------------------------------------------------ */
/* Multilinear(F) — /root/reference/src/poly/multilinear.zig            */
/* ------------------------------------------------------------------ */
int zo_mle_check(uint64_t n, uint32_t *num_vars) { /* init :36-47 */
    if (n == 0) return ZO_ERR_EMPTY_EVALUATIONS;
    if (n & (n - 1)) return ZO_ERR_LENGTH_NOT_POW2;
    if (num_vars) *num_vars = (uint32_t)__builtin_ctzll(n);
    return ZO_OK;
}

uint64_t zo_mle_sum(uint64_t p, const uint64_t *e, uint64_t n) { /* :188-194 */
    uint64_t sum = 0;
    for (uint64_t i = 0; i < n; i++) sum = zo_f_add(p, sum, e[i]);
    return sum;
}

int zo_mle_round_poly(uint64_t p, const uint64_t *e, uint64_t n, uint64_t out[2]) { /* :205-232 */
    if (n < 2) return ZO_ERR_NO_VARIABLES;
    uint64_t half = n / 2, s0 = 0, s1 = 0;
    for (uint64_t i = 0; i < half; i++) {
        s0 = zo_f_add(p, s0, e[i]);
        s1 = zo_f_add(p, s1, e[i + half]);
    }
    out[0] = s0;
    out[1] = zo_f_sub(p, s1, s0);
    return ZO_OK;
}

int zo_mle_partial_eval(uint64_t p, const uint64_t *e, uint64_t n, uint64_t r, uint64_t *out) { /* :154-180 */
    if (n < 2) return ZO_ERR_NO_VARIABLES;
    uint64_t new_len = n / 2;
    for (uint64_t i = 0; i < new_len; i++) {
        uint64_t at0 = e[i], at1 = e[i + new_len];
        uint64_t one_minus_r = zo_f_sub(p, 1 % p, r);
        out[i] = zo_f_add(p, zo_f_mul(p, one_minus_r, at0), zo_f_mul(p, r, at1));
    }
    return ZO_OK;
}

int zo_mle_eval(uint64_t p, const uint64_t *e, uint64_t n, const uint64_t *point, uint32_t npoint, uint64_t *out) { /* :110-144 */
    uint32_t v;
    int rc = zo_mle_check(n, &v);
    if (rc) return rc;
    if (npoint != v) return ZO_ERR_WRONG_NUM_VARS;
    uint64_t(*basis)[2] = malloc(sizeof(uint64_t[2]) * (v ? v : 1));
    if (!basis) return ZO_ERR_OOM;
    for (uint32_t i = 0; i < v; i++) {
        basis[i][0] = zo_f_sub(p, 1 % p, point[i]);
        basis[i][1] = point[i];
    }
    uint64_t result = 0;
    for (uint64_t idx = 0; idx < n; idx++) {
        uint64_t term = e[idx], index = idx;
        for (uint32_t k = 0; k < v; k++) { /* LSB-first: point[k] <-> bit k */
            term = zo_f_mul(p, term, basis[k][index & 1]);
            index >>= 1;
        }
        result = zo_f_add(p, result, term);
    }
    free(basis);
    *out = result;
    return ZO_OK;
}

void zo_mle_add(uint64_t p, const uint64_t *a, const uint64_t *b, uint64_t n, uint64_t *out) { /* :235-250 */
    for (uint64_t i = 0; i < n; i++) out[i] = zo_f_add(p, a[i], b[i]);
}

void zo_mle_scalar_mul(uint64_t p, const uint64_t *a, uint64_t s, uint64_t n, uint64_t *out) { /* :253-264 */
    for (uint64_t i = 0; i < n; i++) out[i] = zo_f_mul(p, a[i], s);
}

/* ------------------------------------------------------------------ */
/* hash.zig                                                            */
/* ------------------------------------------------------------------ */
static void le64(uint64_t v, uint8_t out[8]) {
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)(v >> (8 * i));
}

void zo_transcript_init(zo_transcript *t) { zo_sha3_256_init(&t->h); } /* :261-276 */

void zo_transcript_append_field(zo_transcript *t, uint64_t value) { /* :279-283 */
    uint8_t b[8];
    le64(value, b);
    zo_sha3_256_update(&t->h, b, 8);
}

void zo_transcript_append_bytes(zo_transcript *t, const void *d, size_t n) { zo_sha3_256_update(&t->h, d, n); } /* :293-295 */

uint64_t zo_transcript_challenge(zo_transcript *t, uint64_t p) { /* :301-316 */
    uint8_t digest[32];
    zo_sha3_256_peek(&t->h, digest); /* clone + final */
    uint64_t value = 0;              /* digestToFieldElement :228-242: first 8 bytes LE, then F.init */
    for (int i = 0; i < 8; i++) value |= (uint64_t)digest[i] << (8 * i);
    uint64_t result = zo_f_init(p, value);
    zo_sha3_256_update(&t->h, digest, 32); /* transcript absorbs its own digest */
    return result;
}

void zo_hash_field_element(uint64_t value, uint8_t out[32]) { /* :135-147 */
    uint8_t b[8];
    le64(value, b);
    zo_sha3_256_oneshot(b, 8, out);
}

void zo_merge_hashes(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]) { /* :187-195 */
    uint8_t b[64];
    memcpy(b, l, 32);
    memcpy(b + 32, r, 32);
    zo_sha3_256_oneshot(b, 64, out);
}

/* ------------------------------------------------------------------ */
/* Sumcheck                                                            */
/* ------------------------------------------------------------------ */
uint64_t zo_eval_univariate(uint64_t p, const uint64_t *c, uint32_t n, uint64_t x) { /* sumcheck_protocol.zig:113-123 */
    if (n == 0) return 0;
    uint64_t result = c[n - 1];
    for (uint32_t i = n - 1; i > 0; i--) result = zo_f_add(p, zo_f_mul(p, result, x), c[i - 1]);
    return result;
}

int zo_sumcheck_prove(uint64_t p, const uint64_t *evals, uint64_t n, uint64_t *round_polys, uint64_t *final_point,
                      uint64_t *final_eval, uint64_t *claimed_sum) { /* sumcheck_prover.zig:26-91 */
    uint32_t v;
    int rc = zo_mle_check(n, &v);
    if (rc) return rc;
    if (v == 0) return ZO_ERR_NO_VARIABLES;
    uint64_t claim = zo_mle_sum(p, evals, n); /* :40 */
    if (claimed_sum) *claimed_sum = claim;
    zo_transcript tr;
    zo_transcript_init(&tr); /* State.init sumcheck_protocol.zig:149-164 */
    uint64_t *cur = malloc(n * sizeof(uint64_t)); /* Multilinear.init dupes :47 */
    if (!cur) return ZO_ERR_OOM;
    memcpy(cur, evals, n * sizeof(uint64_t));
    uint64_t len = n;
    for (uint32_t round = 0; round < v; round++) { /* :50-77 */
        uint64_t coeffs[2];
        zo_mle_round_poly(p, cur, len, coeffs);
        round_polys[2 * round] = coeffs[0];
        round_polys[2 * round + 1] = coeffs[1];
        zo_transcript_append_field(&tr, coeffs[0]); /* generateChallenge sumcheck_protocol.zig:176-184 */
        zo_transcript_append_field(&tr, coeffs[1]);
        uint64_t r = zo_transcript_challenge(&tr, p);
        claim = zo_eval_univariate(p, coeffs, 2, r); /* :63-70, value unused by the output */
        final_point[round] = r;
        uint64_t *next = malloc((len / 2) * sizeof(uint64_t)); /* partialEval allocates :162 */
        if (!next) { free(cur); return ZO_ERR_OOM; }
        zo_mle_partial_eval(p, cur, len, r, next);
        free(cur);
        cur = next;
        len /= 2;
    }
    *final_eval = cur[0]; /* :88 */
    free(cur);
    (void)claim;
    return ZO_OK;
}

int zo_sumcheck_prove_interactive(uint64_t p, const uint64_t *evals, uint64_t n, const uint64_t *challenges,
                                  uint32_t n_challenges, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval) {
    /* sumcheck_prover.zig:97-144 */
    uint32_t v;
    int rc = zo_mle_check(n, &v);
    if (rc) return rc;
    if (v == 0) return ZO_ERR_NO_VARIABLES;
    if (n_challenges != v) return ZO_ERR_WRONG_NUM_CHALLENGES;
    uint64_t *cur = malloc(n * sizeof(uint64_t));
    if (!cur) return ZO_ERR_OOM;
    memcpy(cur, evals, n * sizeof(uint64_t));
    uint64_t len = n;
    for (uint32_t round = 0; round < v; round++) {
        zo_mle_round_poly(p, cur, len, &round_polys[2 * round]);
        uint64_t *next = malloc((len / 2) * sizeof(uint64_t));
        if (!next) { free(cur); return ZO_ERR_OOM; }
        zo_mle_partial_eval(p, cur, len, challenges[round], next);
        free(cur);
        cur = next;
        len /= 2;
    }
    for (uint32_t i = 0; i < v; i++) final_point[i] = challenges[i];
    *final_eval = cur[0];
    free(cur);
    return ZO_OK;
}

void zo_sumcheck_proof_to_bytes(uint32_t v, const uint64_t *round_polys, const uint64_t *final_point, uint64_t final_eval,
                                uint8_t *out) { /* sumcheck_protocol.zig:76-109 */
    size_t off = 0;
    le64(v, out + off); off += 8;
    for (uint32_t i = 0; i < 2 * v; i++) { le64(round_polys[i], out + off); off += 8; }
    for (uint32_t i = 0; i < v; i++) { le64(final_point[i], out + off); off += 8; }
    le64(final_eval, out + off);
}

int zo_sumcheck_verify_rounds(uint64_t p, uint32_t v, uint32_t ncoef, const uint64_t *round_polys, uint64_t claimed_sum,
                              int *is_valid, uint64_t *final_claim) { /* sumcheck_verifier.zig:172-202 */
    zo_transcript tr;
    zo_transcript_init(&tr);
    uint64_t claim = claimed_sum;
    for (uint32_t round = 0; round < v; round++) {
        const uint64_t *rp = round_polys + (size_t)round * ncoef;
        uint64_t e0 = zo_eval_univariate(p, rp, ncoef, 0);
        uint64_t e1 = zo_eval_univariate(p, rp, ncoef, 1 % p);
        if (zo_f_add(p, e0, e1) != claim) {
            *is_valid = 0;
            *final_claim = 0;
            return ZO_OK;
        }
        for (uint32_t k = 0; k < ncoef; k++) zo_transcript_append_field(&tr, rp[k]);
        uint64_t r = zo_transcript_challenge(&tr, p);
        claim = zo_eval_univariate(p, rp, ncoef, r);
    }
    *is_valid = 1;
    *final_claim = claim;
    return ZO_OK;
}

/* ---- extension: product of d multilinears (no reference behaviour for d > 1) ---- */
/* coefficients [a0..ad] of g(X) = sum_i prod_k (lo_k[i] + (hi_k[i] - lo_k[i]) X) over MSB-first pairs (i, i + n/2) */
int zo_prod_round_coeffs(uint64_t p, const uint64_t *const *polys, uint32_t d, uint64_t n, uint64_t *out) {
    if (n < 2 || d == 0 || d > 3) return ZO_ERR_NO_VARIABLES;
    uint64_t half = n / 2;
    uint64_t acc[4] = {0, 0, 0, 0};
    for (uint64_t i = 0; i < half; i++) {
        uint64_t c[4] = {1 % p, 0, 0, 0};
        for (uint32_t k = 0; k < d; k++) {
            uint64_t lo = polys[k][i], df = zo_f_sub(p, polys[k][i + half], lo);
            uint64_t nc[4] = {0, 0, 0, 0};
            for (uint32_t j = 0; j <= k; j++) {
                nc[j] = zo_f_add(p, nc[j], zo_f_mul(p, c[j], lo));
                nc[j + 1] = zo_f_add(p, nc[j + 1], zo_f_mul(p, c[j], df));
            }
            memcpy(c, nc, sizeof(c));
        }
        for (uint32_t j = 0; j <= d; j++) acc[j] = zo_f_add(p, acc[j], c[j]);
    }
    for (uint32_t j = 0; j <= d; j++) out[j] = acc[j];
    return ZO_OK;
}

int zo_prodcheck_prove(uint64_t p, const uint64_t *const *polys, uint32_t d, uint64_t n, uint64_t *round_polys,
                       uint64_t *final_point, uint64_t *final_evals, uint64_t *claimed_sum) {
    uint32_t v;
    int rc = zo_mle_check(n, &v);
    if (rc) return rc;
    if (v == 0 || d == 0 || d > 3) return ZO_ERR_NO_VARIABLES;
    uint64_t *cur[3] = {0, 0, 0};
    for (uint32_t k = 0; k < d; k++) {
        cur[k] = malloc(n * sizeof(uint64_t));
        if (!cur[k]) return ZO_ERR_OOM;
        memcpy(cur[k], polys[k], n * sizeof(uint64_t));
    }
    if (claimed_sum) {
        uint64_t s = 0;
        for (uint64_t i = 0; i < n; i++) {
            uint64_t t = cur[0][i];
            for (uint32_t k = 1; k < d; k++) t = zo_f_mul(p, t, cur[k][i]);
            s = zo_f_add(p, s, t);
        }
        *claimed_sum = s;
    }
    zo_transcript tr;
    zo_transcript_init(&tr);
    uint64_t len = n;
    for (uint32_t round = 0; round < v; round++) {
        uint64_t half = len / 2;
        uint64_t acc[4] = {0, 0, 0, 0};
        zo_prod_round_coeffs(p, (const uint64_t *const *)cur, d, len, acc);
        for (uint32_t j = 0; j <= d; j++) {
            round_polys[(size_t)round * (d + 1) + j] = acc[j];
            zo_transcript_append_field(&tr, acc[j]);
        }
        uint64_t r = zo_transcript_challenge(&tr, p);
        final_point[round] = r;
        for (uint32_t k = 0; k < d; k++) {
            uint64_t *next = malloc(half * sizeof(uint64_t));
            if (!next) return ZO_ERR_OOM;
            zo_mle_partial_eval(p, cur[k], len, r, next);
            free(cur[k]);
            cur[k] = next;
        }
        len = half;
    }
    for (uint32_t k = 0; k < d; k++) {
        final_evals[k] = cur[k][0];
        free(cur[k]);
    }
    return ZO_OK;
}

/* ------------------------------------------------------------------ */
/* SimpleMerkleTree — /root/reference/src/commitments/merkle_tree.zig   */
/* ------------------------------------------------------------------ */
uint64_t zo_ceil_pow2(uint64_t n) {
    uint64_t r = 1;
    while (r < n) r <<= 1;
    return r;
}

static int compute_root(const uint8_t *hashes, uint64_t count, uint8_t root[32]) { /* computeRoot :380-400 */
    if (count == 1) {
        memcpy(root, hashes, 32);
        return ZO_OK;
    }
    uint8_t *cur = malloc(count * 32);
    if (!cur) return ZO_ERR_OOM;
    memcpy(cur, hashes, count * 32);
    while (count > 1) {
        uint64_t next_size = count / 2;
        uint8_t *next = malloc(next_size * 32);
        if (!next) { free(cur); return ZO_ERR_OOM; }
        for (uint64_t i = 0; i < next_size; i++) zo_merge_hashes(cur + 64 * i, cur + 64 * i + 32, next + 32 * i);
        free(cur);
        cur = next;
        count = next_size;
    }
    memcpy(root, cur, 32);
    free(cur);
    return ZO_OK;
}

int zo_merkle_build(const uint64_t *values, uint64_t n, uint8_t *leaf_hashes, uint8_t root[32], uint32_t *height) { /* :283-318 */
    if (n == 0) return ZO_ERR_EMPTY_VALUES;
    uint64_t padded = zo_ceil_pow2(n);
    if (height) *height = (uint32_t)__builtin_ctzll(padded);
    uint8_t *lh = leaf_hashes ? leaf_hashes : malloc(padded * 32);
    if (!lh) return ZO_ERR_OOM;
    for (uint64_t i = 0; i < n; i++) zo_hash_field_element(values[i], lh + 32 * i);
    uint8_t zero_hash[32];
    zo_hash_field_element(0, zero_hash);
    for (uint64_t i = n; i < padded; i++) memcpy(lh + 32 * i, zero_hash, 32);
    int rc = compute_root(lh, padded, root);
    if (!leaf_hashes) free(lh);
    return rc;
}

int zo_merkle_open(const uint64_t *values, uint64_t n, const uint8_t *leaf_hashes, uint64_t index, uint8_t *siblings,
                   uint8_t *dirs, uint64_t *value) { /* :324-360 */
    if (index >= n) return ZO_ERR_INDEX_OUT_OF_BOUNDS;
    uint64_t count = zo_ceil_pow2(n);
    uint32_t height = (uint32_t)__builtin_ctzll(count);
    uint64_t cur_index = index;
    uint8_t *cur = malloc(count * 32);
    if (!cur) return ZO_ERR_OOM;
    memcpy(cur, leaf_hashes, count * 32);
    for (uint32_t level = 0; level < height; level++) {
        int is_right = (cur_index % 2) == 1;
        uint64_t sib = is_right ? cur_index - 1 : cur_index + 1;
        memcpy(siblings + 32 * level, cur + 32 * sib, 32);
        dirs[level] = (uint8_t)is_right;
        uint64_t next_size = count / 2;
        uint8_t *next = malloc(next_size * 32);
        if (!next) { free(cur); return ZO_ERR_OOM; }
        for (uint64_t i = 0; i < next_size; i++) zo_merge_hashes(cur + 64 * i, cur + 64 * i + 32, next + 32 * i);
        free(cur);
        cur = next;
        count = next_size;
        cur_index /= 2;
    }
    free(cur);
    if (value) *value = values[index];
    return ZO_OK;
}

/* `k` openings with ONE recomputation of the levels: exactly zo_merkle_open's walk (merkle_tree.zig:335-353) for every
 * index, sharing the level arrays — a test convenience for large trees (the reference recomputes all levels per open).
 * siblings: k x height x 32 bytes, dirs: k x height, values: k. */
int zo_merkle_open_many(const uint64_t *values, uint64_t n, const uint8_t *leaf_hashes, const uint64_t *indices, uint32_t k,
                        uint8_t *siblings, uint8_t *dirs, uint64_t *out_values) {
    for (uint32_t j = 0; j < k; j++)
        if (indices[j] >= n) return ZO_ERR_INDEX_OUT_OF_BOUNDS;
    uint64_t count = zo_ceil_pow2(n);
    uint32_t height = (uint32_t)__builtin_ctzll(count);
    uint8_t *cur = malloc(count * 32);
    if (!cur) return ZO_ERR_OOM;
    memcpy(cur, leaf_hashes, count * 32);
    for (uint32_t level = 0; level < height; level++) {
        for (uint32_t j = 0; j < k; j++) {
            uint64_t ci = indices[j] >> level;
            int is_right = (ci % 2) == 1;
            uint64_t sib = is_right ? ci - 1 : ci + 1;
            memcpy(siblings + ((size_t)j * height + level) * 32, cur + 32 * sib, 32);
            dirs[(size_t)j * height + level] = (uint8_t)is_right;
        }
        uint64_t next_size = count / 2;
        uint8_t *next = malloc(next_size * 32);
        if (!next) { free(cur); return ZO_ERR_OOM; }
        for (uint64_t i = 0; i < next_size; i++) zo_merge_hashes(cur + 64 * i, cur + 64 * i + 32, next + 32 * i);
        free(cur);
        cur = next;
        count = next_size;
    }
    free(cur);
    for (uint32_t j = 0; j < k; j++) out_values[j] = values[indices[j]];
    return ZO_OK;
}

int zo_merkle_verify(const uint8_t root[32], uint64_t value, const uint8_t *siblings, const uint8_t *dirs, uint32_t height) {
    /* :362-373 */
    uint8_t cur[32], nxt[32];
    zo_hash_field_element(value, cur);
    for (uint32_t l = 0; l < height; l++) {
        if (dirs[l]) zo_merge_hashes(siblings + 32 * l, cur, nxt);
        else zo_merge_hashes(cur, siblings + 32 * l, nxt);
        memcpy(cur, nxt, 32);
    }
    return memcmp(cur, root, 32) == 0;
}

/* ------------------------------------------------------------------ */
/* CommitmentScheme — polynomial_commit.zig                            */
/* ------------------------------------------------------------------ */
uint64_t zo_point_to_index(const uint64_t *point, uint32_t npoint) { /* :178-183 */
    if (npoint == 0) return 0;
    return point[0] % ((uint64_t)1 << npoint);
}

int zo_commit_open(uint64_t p, const uint64_t *evals, uint64_t n, const uint8_t *leaf_hashes, const uint64_t *point,
                   uint32_t npoint, uint64_t *value, uint64_t *leaf_index, uint64_t *leaf_value, uint8_t *siblings,
                   uint8_t *dirs) { /* open :86-115 */
    uint32_t v;
    int rc = zo_mle_check(n, &v);
    if (rc) return rc;
    if (npoint != v) return ZO_ERR_POINT_DIM_MISMATCH;
    rc = zo_mle_eval(p, evals, n, point, npoint, value);
    if (rc) return rc;
    uint64_t index = zo_point_to_index(point, npoint);
    *leaf_index = index;
    return zo_merkle_open(evals, n, leaf_hashes, index, siblings, dirs, leaf_value);
}

/* ------------------------------------------------------------------ */
/* Prover.generateCommitments — /root/reference/src/prover/prover.zig:366-467 */
/* ------------------------------------------------------------------ */
int zo_generate_commitments(uint64_t p, zo_transcript *tr, const uint64_t *const *polys, uint32_t count, uint64_t n,
                            uint8_t *roots, uint64_t *points, uint64_t *values, uint64_t *leaf_indices, uint64_t *leaf_values,
                            uint8_t *siblings, uint8_t *dirs) {
    uint32_t v;
    int rc = zo_mle_check(n, &v);
    if (rc) return rc;
    uint8_t **lh = calloc(count, sizeof(uint8_t *));
    if (!lh) return ZO_ERR_OOM;
    /* PHASE 1 (:405-410): commit every polynomial */
    for (uint32_t i = 0; i < count && rc == ZO_OK; i++) {
        lh[i] = malloc(n * 32);
        if (!lh[i]) rc = ZO_ERR_OOM;
        else rc = zo_merkle_build(polys[i], n, lh[i], roots + 32 * (size_t)i, NULL);
    }
    if (rc == ZO_OK) {
        /* PHASE 2 (:413-416): bind all commitments */
        zo_transcript_append_bytes(tr, "POLY_COMMITMENTS", 16);
        for (uint32_t i = 0; i < count; i++) zo_transcript_append_bytes(tr, roots + 32 * (size_t)i, 32);
        /* PHASE 3 (:420-443): challenges, evaluation, opening */
        for (uint32_t i = 0; i < count && rc == ZO_OK; i++) {
            uint64_t *pt = points + (size_t)i * v;
            for (uint32_t j = 0; j < v; j++) pt[j] = zo_transcript_challenge(tr, p);
            rc = zo_mle_eval(p, polys[i], n, pt, v, &values[i]); /* :427 */
            if (rc) break;
            uint64_t v2;
            rc = zo_commit_open(p, polys[i], n, lh[i], pt, v, &v2, &leaf_indices[i], &leaf_values[i],
                                siblings + (size_t)i * v * 32, dirs + (size_t)i * v); /* :431 (evaluates again) */
        }
        /* PHASE 4 (:463-466): bind the opening claims */
        if (rc == ZO_OK) {
            zo_transcript_append_bytes(tr, "OPENING_CLAIMS", 14);
            for (uint32_t i = 0; i < count; i++) zo_transcript_append_field(tr, values[i]);
        }
    }
    for (uint32_t i = 0; i < count; i++) free(lh[i]);
    free(lh);
    return rc;
}

/* ------------------------------------------------------------------ */
/* WitnessGenerator.generate — /root/reference/src/constraints/witness.zig:29-270, on a column (SoA) view of the trace:
 * cols[c][i] = raw u64 of column c at step i, in the order of prover.zig:376-390 (pc, x0..x31, opcode, rd, rs1, rs2,
 * funct3, funct7, imm, mem.address, mem.value, mem.is_read). The first n_hold columns (pc + registers) pad by repeating
 * the last value (:80-87, :116-123), the others pad with zero (:174-182, :249-253). F.init = value mod p. */
/* ------------------------------------------------------------------ */
uint64_t zo_witness_pack(uint64_t p, const uint64_t *cols, uint64_t num_steps, uint32_t n_cols, uint32_t n_hold, uint64_t *out) {
    uint32_t num_vars = 0;
    while (((uint64_t)1 << num_vars) < num_steps) num_vars++; /* log2_int_ceil, 0 for 0 or 1 steps (:36-39) */
    uint64_t padded = (uint64_t)1 << num_vars;
    for (uint32_t c = 0; c < n_cols; c++) {
        const uint64_t *col = cols + (size_t)c * num_steps;
        uint64_t *o = out + (size_t)c * padded;
        for (uint64_t i = 0; i < num_steps; i++) o[i] = zo_f_init(p, col[i]);
        uint64_t fill = (c < n_hold && num_steps > 0) ? zo_f_init(p, col[num_steps - 1]) : 0;
        for (uint64_t i = num_steps; i < padded; i++) o[i] = fill;
    }
    return padded;
}

/* ------------------------------------------------------------------ */
/* Prover.prove after the VM + BinarySerializer.serialize + Verifier.verify                                          */
/* /root/reference/src/prover/prover.zig:73-226, 229-362, 514-559; src/prover/proof.zig:194-261;                      */
/* src/prover/serialization.zig:70-97, 134-245, 296-344, 374-429; src/verifier/verifier.zig:49-294                    */
/* ------------------------------------------------------------------ */
static void put32(uint8_t **w, uint32_t v) { for (int i = 0; i < 4; i++) *(*w)++ = (uint8_t)(v >> (8 * i)); }
static void put64(uint8_t **w, uint64_t v) { for (int i = 0; i < 8; i++) *(*w)++ = (uint8_t)(v >> (8 * i)); }
static uint32_t get32(const uint8_t **r) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)(*(*r)++) << (8 * i); return v; }
static uint64_t get64(const uint8_t **r) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)(*(*r)++) << (8 * i); return v; }

static int opcode_has_table(uint64_t opcode) { /* instruction_table.zig:243-275 via state.zig:151 */
    return opcode == 0x33 || opcode == 0x13 || opcode == 0x03 || opcode == 0x23 || opcode == 0x63;
}

uint64_t zo_count_lookups(const uint64_t *opcode_col, uint64_t num_steps) { /* builder.zig:253-267 */
    uint64_t c = 0;
    for (uint64_t i = 0; i < num_steps; i++) c += opcode_has_table(opcode_col[i]);
    return c;
}

static uint32_t log2_ceil(uint64_t n) {
    uint32_t v = 0;
    while (((uint64_t)1 << v) < n) v++;
    return v;
}

size_t zo_proof_exact_size(uint64_t num_steps, uint32_t n_init, uint32_t n_out, uint64_t n_lookups) {
    uint32_t v = log2_ceil(num_steps);
    return 32 + (32 + 8 + 8 + 4 + 8 * (size_t)n_init + 4 + 8 * 32 + 8 + 4 + 8 * (size_t)n_out) + ((size_t)v * 4 * 8 + (size_t)v * 8 + 8) + 4 +
           (size_t)n_lookups * 24 + 43 * (32 + (size_t)v * 8 + 8 + 8 + 8 + 8 + 4 + (size_t)v * 33);
}

size_t zo_proof_estimated_size(uint64_t num_steps, uint32_t n_init, uint64_t n_lookups) { /* serialization.zig:134-173 */
    uint32_t v = log2_ceil(num_steps);
    return 32 + (32 + 8 + 8 + 4 + 4) + 8 * (size_t)n_init + 8 * 32 + ((size_t)v * 4 * 8 + (size_t)v * 8 + 8) + 4 + (size_t)n_lookups * (4 + 8 + 8) +
           43 * (32 + (size_t)v * 8 + 8 + 32 * 20);
}

int zo_prove_from_trace(uint64_t p, const uint8_t *program, size_t program_len, uint64_t entry_pc, const uint64_t *initial_regs,
                        uint32_t n_init, const uint64_t *cols, uint64_t num_steps, uint64_t final_pc, const uint64_t *final_regs,
                        const uint64_t *outputs, uint32_t n_out, int compat_buffer, uint8_t *out, size_t out_cap, size_t *out_len) {
    if (num_steps == 0) return ZO_ERR_EMPTY_TRACE; /* prover.zig:144-146 */
    const uint32_t v = log2_ceil(num_steps);
    const uint64_t padded = (uint64_t)1 << v;
    const uint64_t n_lookups = zo_count_lookups(cols + 33 * num_steps, num_steps);
    const size_t exact = zo_proof_exact_size(num_steps, n_init, n_out, n_lookups);
    if (compat_buffer && exact > zo_proof_estimated_size(num_steps, n_init, n_lookups)) return ZO_ERR_NO_SPACE_LEFT; /* :72-73 */
    if (out_len) *out_len = exact;
    if (!out || out_cap < exact) return ZO_ERR_OOM;
    zo_transcript tr;
    zo_transcript_init(&tr); /* :91 */
    uint8_t program_hash[32];
    zo_sha256_oneshot(program, program_len, program_hash); /* :98-100 */
    zo_transcript_append_bytes(&tr, program_hash, 32);
    zo_transcript_append_field(&tr, zo_f_init(p, entry_pc)); /* :103 */
    for (uint32_t i = 0; i < n_init; i++) zo_transcript_append_field(&tr, zo_f_init(p, initial_regs[i])); /* :106-110 */
    /* witness (witness.zig:29-270) */
    uint64_t *w = malloc(43 * padded * sizeof(uint64_t));
    if (!w) return ZO_ERR_OOM;
    zo_witness_pack(p, cols, num_steps, 43, 33, w);
    uint8_t *o = out;
    /* header serialization.zig:175-182 */
    memcpy(o, "ZIGZ", 4); o += 4;
    put32(&o, 1); put64(&o, p); put64(&o, num_steps); put32(&o, v); put32(&o, 0);
    /* public IO :209-245 (packagePublicIO prover.zig:514-559) */
    memcpy(o, program_hash, 32); o += 32;
    put64(&o, entry_pc); put64(&o, final_pc);
    put32(&o, n_init);
    for (uint32_t i = 0; i < n_init; i++) put64(&o, initial_regs[i]);
    put32(&o, 32);
    for (int i = 0; i < 32; i++) put64(&o, final_regs[i]);
    put64(&o, num_steps);
    put32(&o, n_out);
    for (uint32_t i = 0; i < n_out; i++) put64(&o, outputs[i]);
    /* constraint sumcheck placeholder prover.zig:229-289 */
    zo_transcript_append_bytes(&tr, "SUMCHECK_BEGIN", 14);
    zo_transcript_append_field(&tr, zo_f_init(p, num_steps));
    zo_transcript_append_field(&tr, zo_f_init(p, v));
    uint64_t chal[64];
    for (uint32_t r = 0; r < v; r++) {
        for (int k = 0; k < 4; k++) zo_transcript_append_field(&tr, 0);
        chal[r] = zo_transcript_challenge(&tr, p);
    }
    for (uint32_t r = 0; r < v; r++) for (int k = 0; k < 4; k++) put64(&o, 0); /* serialization.zig:296-311 */
    for (uint32_t r = 0; r < v; r++) put64(&o, chal[r]);
    put64(&o, 0);
    /* Lasso placeholders prover.zig:292-362: one 0-round proof per lookup constraint */
    zo_transcript_append_bytes(&tr, "LASSO_BEGIN", 11);
    put32(&o, (uint32_t)n_lookups);
    for (uint64_t k = 0; k < n_lookups; k++) {
        zo_transcript_append_bytes(&tr, "LASSO_TABLE", 11);
        zo_transcript_append_field(&tr, zo_f_init(p, (uint32_t)k));
        put32(&o, (uint32_t)k); put64(&o, 1); put32(&o, 0); put64(&o, 0); /* :333-344: id, num_lookups, num_vars, final_eval */
    }
    /* commitments prover.zig:366-467 */
    const uint64_t *polys[43];
    for (int i = 0; i < 43; i++) polys[i] = w + (size_t)i * padded;
    uint8_t *roots = malloc(43 * 32), *sib = malloc(43 * (size_t)(v ? v : 1) * 32), *dirs = malloc(43 * (size_t)(v ? v : 1));
    uint64_t *pts = malloc(43 * (size_t)(v ? v : 1) * 8), vals[43], li[43], lv[43];
    int rc = (roots && sib && dirs && pts) ? zo_generate_commitments(p, &tr, polys, 43, padded, roots, pts, vals, li, lv, sib, dirs) : ZO_ERR_OOM;
    if (rc == ZO_OK) {
        for (int i = 0; i < 43; i++) { /* serialization.zig:374-429 */
            memcpy(o, roots + 32 * i, 32); o += 32;
            for (uint32_t j = 0; j < v; j++) put64(&o, pts[(size_t)i * v + j]);
            put64(&o, vals[i]);
            put64(&o, vals[i]); /* OpeningProof.value (the same evaluation, polynomial_commit.zig:97) */
            put64(&o, li[i]); put64(&o, lv[i]); put32(&o, v);
            memcpy(o, sib + (size_t)i * v * 32, (size_t)v * 32); o += (size_t)v * 32;
            for (uint32_t j = 0; j < v; j++) *o++ = dirs[(size_t)i * v + j] ? 1 : 0;
        }
    }
    free(w); free(roots); free(sib); free(dirs); free(pts);
    if (rc == ZO_OK && (size_t)(o - out) != exact) rc = ZO_ERR_OOM; /* internal consistency */
    return rc;
}

/* Verifier.verify on serialized bytes: 0 Accept, 1 RejectInvalidSumcheck, 2 RejectInvalidLookup, 3 RejectInvalidCommitment;
 * negative: ZO_ERR_PROGRAM_HASH_MISMATCH / ZO_ERR_BAD_PROOF (deserialize errors) */
int zo_verify_proof(uint64_t p, const uint8_t *proof, size_t len, const uint8_t *program, size_t program_len, int *result) {
    const uint8_t *r = proof, *end = proof + len;
#define NEED(nbytes) do { if ((size_t)(end - r) < (size_t)(nbytes)) return ZO_ERR_BAD_PROOF; } while (0)
    NEED(32);
    if (memcmp(r, "ZIGZ", 4)) return ZO_ERR_BAD_PROOF;
    r += 4;
    if (get32(&r) != 1) return ZO_ERR_BAD_PROOF;
    if (get64(&r) != p) return ZO_ERR_BAD_PROOF;
    uint64_t num_steps = get64(&r);
    uint32_t v = get32(&r);
    (void)get32(&r);
    if (v != log2_ceil(num_steps) || v > 63) return ZO_ERR_BAD_PROOF; /* Proof.init derives num_vars from num_steps (serialization.zig:113) */
    NEED(32 + 16 + 4);
    uint8_t hash[32], ph[32];
    memcpy(ph, r, 32); r += 32;
    zo_sha256_oneshot(program, program_len, hash);
    (void)get64(&r); (void)get64(&r);
    uint32_t n = get32(&r); NEED(8 * (size_t)n + 4); r += 8 * (size_t)n;
    n = get32(&r); NEED(8 * (size_t)n + 12); r += 8 * (size_t)n;
    (void)get64(&r);
    n = get32(&r); NEED(8 * (size_t)n); r += 8 * (size_t)n;
    /* deserialize (serialization.zig:98-127) reads EVERY section before verify runs: structural errors anywhere come
     * first, then the program-hash error (verifier.zig:101-107), then the verdicts in the verifier's order */
    /* constraint sumcheck: only round 0 is checked (verifier.zig:209-214) */
    NEED((size_t)v * 40 + 8);
    uint64_t g0 = 0, g1 = 0;
    for (uint32_t rd = 0; rd < v; rd++)
        for (int k = 0; k < 4; k++) {
            uint64_t c = zo_f_init(p, get64(&r));
            if (rd == 0) { if (k == 0) g0 = c; g1 = zo_f_add(p, g1, c); }
        }
    r += (size_t)v * 8;
    uint64_t final_eval = zo_f_init(p, get64(&r));
    *result = 0;
    if (v > 0 && zo_f_add(p, g0, g1) != final_eval) *result = 1;
    NEED(4);
    uint32_t n_lasso = get32(&r);
    for (uint32_t k = 0; k < n_lasso; k++) {
        NEED(16);
        (void)get32(&r); (void)get64(&r);
        uint32_t lv_ = get32(&r);
        NEED((size_t)lv_ * 32 + 8);
        uint64_t a0 = 0, a1 = 0;
        for (uint32_t rd = 0; rd < lv_; rd++)
            for (int c3 = 0; c3 < 3; c3++) { /* LassoProof multiset_proof has 3 coefficients per round (proof.zig:102-140) */
                uint64_t c = zo_f_init(p, get64(&r));
                if (rd == 0) { if (c3 == 0) a0 = c; a1 = zo_f_add(p, a1, c); }
            }
        r += (size_t)lv_ * 8;
        uint64_t fe = zo_f_init(p, get64(&r));
        if (lv_ > 0 && zo_f_add(p, a0, a1) != fe && *result == 0) *result = 2;
    }
    for (int i = 0; i < 43; i++) {
        NEED(32 + (size_t)v * 8 + 8 + 28);
        const uint8_t *root = r; r += 32;
        r += (size_t)v * 8;
        uint64_t value = zo_f_init(p, get64(&r)), pvalue = zo_f_init(p, get64(&r));
        (void)get64(&r);
        uint64_t leaf = zo_f_init(p, get64(&r));
        uint32_t plen = get32(&r);
        NEED((size_t)plen * 33);
        const uint8_t *sibs = r; r += (size_t)plen * 32;
        const uint8_t *dirs = r; r += plen;
        /* verifyOpening verifier.zig:270-294: value == proof.value, point.len == num_vars, Merkle path */
        if (*result == 0 && (value != pvalue || !zo_merkle_verify(root, leaf, sibs, dirs, plen))) *result = 3;
    }
#undef NEED
    if (memcmp(hash, ph, 32)) { *result = 0; return ZO_ERR_PROGRAM_HASH_MISMATCH; } /* verifier.zig:101-107 */
    return ZO_OK;
}

/* ------------------------------------------------------------------ */
/* Lasso — lasso_prover.zig, table_builder.zig                          */
/* ------------------------------------------------------------------ */
uint64_t zo_lasso_hash_row(uint64_t p, const uint64_t *row, uint32_t arity) { /* hashEntry/hashQuery :208-239 */
    uint64_t h = 0;
    for (uint32_t k = 0; k < arity; k++) {
        h ^= row[k];
        uint8_t b[8];
        le64(h, b);
        h = zo_xxh3_64_small(b, 8, 0);
    }
    return zo_f_init(p, h % p);
}

void zo_lasso_commit_poly(const uint64_t *evals, uint64_t n, uint8_t out[32]) { /* commitToPolynomial :242-252 */
    zo_sha3_256 c;
    zo_sha3_256_init(&c);
    for (uint64_t i = 0; i < n; i++) {
        uint8_t b[8];
        le64(evals[i], b);
        zo_sha3_256_update(&c, b, 8);
    }
    zo_sha3_256_peek(&c, out);
}

void zo_build_table(uint64_t p, int op, uint32_t bits, uint64_t *rows) { /* table_builder.zig:126-213 */
    uint64_t max_val = (uint64_t)1 << bits, idx = 0;
    for (uint64_t a = 0; a < max_val; a++)
        for (uint64_t b = 0; b < max_val; b++) {
            uint64_t r = op == ZO_TABLE_ADD ? (a + b) % max_val : op == ZO_TABLE_XOR ? (a ^ b) : (a & b);
            rows[3 * idx] = zo_f_init(p, a);
            rows[3 * idx + 1] = zo_f_init(p, b);
            rows[3 * idx + 2] = zo_f_init(p, r);
            idx++;
        }
}

int zo_lasso_prove(uint64_t p, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows, uint64_t n_queries,
                   uint32_t arity, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                   uint8_t query_commitment[32], uint8_t table_commitment[32]) { /* prove :103-173 */
    if (n_queries == 0) return ZO_ERR_NO_QUERIES;
    uint64_t *table_evals = malloc((n_table ? n_table : 1) * sizeof(uint64_t));
    if (!table_evals) return ZO_ERR_OOM;
    for (uint64_t i = 0; i < n_table; i++) table_evals[i] = zo_lasso_hash_row(p, table_rows + (size_t)i * arity, arity);
    int rc = zo_mle_check(n_table, NULL); /* Multilinear.init(table_evals) :124 */
    if (rc) { free(table_evals); return rc; }
    uint64_t padded = zo_ceil_pow2(n_queries);
    uint64_t *query_evals = malloc(padded * sizeof(uint64_t));
    if (!query_evals) { free(table_evals); return ZO_ERR_OOM; }
    for (uint64_t j = 0; j < n_queries; j++) query_evals[j] = zo_lasso_hash_row(p, query_rows + (size_t)j * arity, arity);
    for (uint64_t j = n_queries; j < padded; j++) query_evals[j] = 0; /* :140-142 */
    (void)zo_mle_sum(p, query_evals, padded);                          /* :154, result discarded */
    uint32_t v = (uint32_t)__builtin_ctzll(padded);
    if (num_vars) *num_vars = v;
    rc = zo_sumcheck_prove(p, query_evals, padded, round_polys, final_point, final_eval, NULL); /* :160 */
    if (rc == ZO_OK) {
        zo_lasso_commit_poly(query_evals, padded, query_commitment); /* :163 */
        zo_lasso_commit_poly(table_evals, n_table, table_commitment); /* :164 */
    }
    free(query_evals);
    free(table_evals);
    return rc;
}

int zo_lasso_prove_with_mapping(uint64_t p, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows,
                                uint64_t n_queries, const uint64_t *mapping, uint64_t n_mapping, uint32_t arity,
                                uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                                uint8_t query_commitment[32], uint8_t table_commitment[32]) { /* :179-205 */
    if (n_queries != n_mapping) return ZO_ERR_MAPPING_LEN_MISMATCH;
    for (uint64_t j = 0; j < n_queries; j++) {
        if (mapping[j] >= n_table) return ZO_ERR_INVALID_MAPPING;
        if (memcmp(query_rows + (size_t)j * arity, table_rows + (size_t)mapping[j] * arity, arity * sizeof(uint64_t)) != 0)
            return ZO_ERR_QUERY_TABLE_MISMATCH; /* entriesMatch :255-268 */
    }
    return zo_lasso_prove(p, table_rows, n_table, query_rows, n_queries, arity, round_polys, final_point, final_eval, num_vars,
                          query_commitment, table_commitment);
}

/* ------------------------------------------------------------------ */
/* synthetic inputs                                                    */
/* ------------------------------------------------------------------ */
uint64_t zo_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

void zo_fill_synthetic(uint64_t p, uint64_t seed, uint64_t start, uint64_t n, uint64_t *out) {
    for (uint64_t i = 0; i < n; i++) out[i] = zo_splitmix64(seed + start + i) % p;
}
