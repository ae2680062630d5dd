/* oracle/zigz_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in plain C, of the zigz proving hot path.
 * Nothing under zigz_b200/ may include, link or call this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: the reference (Zig 0.15.2 + un-vendored hash-zig) cannot be
 * built in this image and ships no digest / proof golden vectors, so the
 * hash-dependent outputs are pinned by construction only:
 *   - the numeric facts the reference's own unit tests assert (tests/test_oracle_reference_facts.py)
 *   - SHA3-256 / SHA-256 / XXH3-64 against hashlib and the xxhash package
 *   - the seed vectors recorded in SURVEY.md §8(c)
 * => "parity unpinned" against a real `zig build` for digests; arithmetic is pinned.
 *
 * All field elements cross this interface as canonical uint64_t in [0, p),
 * the reference's `struct { value: u64 }` (src/core/field.zig:26-27).
 */
#ifndef ZIGZ_ORACLE_H
#define ZIGZ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "zo_hash.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ZO_BABYBEAR_P 2013265921ULL /* src/core/field_presets.zig:19 */

/* error codes = the reference's Zig error names */
enum {
    ZO_OK = 0,
    ZO_ERR_EMPTY_EVALUATIONS = -1,      /* multilinear.zig:38 */
    ZO_ERR_LENGTH_NOT_POW2 = -2,        /* multilinear.zig:43 */
    ZO_ERR_WRONG_NUM_VARS = -3,         /* multilinear.zig:112 */
    ZO_ERR_NO_VARIABLES = -4,           /* multilinear.zig:156,207 sumcheck_prover.zig:31 */
    ZO_ERR_EMPTY_VALUES = -5,           /* merkle_tree.zig:284 */
    ZO_ERR_INDEX_OUT_OF_BOUNDS = -6,    /* merkle_tree.zig:325 */
    ZO_ERR_POINT_DIM_MISMATCH = -7,     /* polynomial_commit.zig:93 */
    ZO_ERR_NO_QUERIES = -8,             /* lasso_prover.zig:109 */
    ZO_ERR_MAPPING_LEN_MISMATCH = -9,   /* lasso_prover.zig:186 */
    ZO_ERR_INVALID_MAPPING = -10,       /* lasso_prover.zig:192 */
    ZO_ERR_QUERY_TABLE_MISMATCH = -11,  /* lasso_prover.zig:199 */
    ZO_ERR_WRONG_NUM_CHALLENGES = -12,  /* sumcheck_prover.zig:105 */
    ZO_ERR_EMPTY_TRACE = -14,           /* prover.zig:145 */
    ZO_ERR_NO_SPACE_LEFT = -15,         /* serialization.zig:72-73 fixed buffer */
    ZO_ERR_PROGRAM_HASH_MISMATCH = -16, /* verifier.zig:105 */
    ZO_ERR_BAD_PROOF = -17,             /* deserialize failures */
    ZO_ERR_OOM = -100
};

/* ---- Field(u64, p): src/core/field.zig ---- */
uint64_t zo_f_init(uint64_t p, uint64_t v);             /* :36-38 */
uint64_t zo_f_add(uint64_t p, uint64_t a, uint64_t b);  /* :73-88 */
uint64_t zo_f_sub(uint64_t p, uint64_t a, uint64_t b);  /* :91-98 */
uint64_t zo_f_neg(uint64_t p, uint64_t a);              /* :101-106 */
uint64_t zo_f_mul(uint64_t p, uint64_t a, uint64_t b);  /* :112-147 */
uint64_t zo_f_pow(uint64_t p, uint64_t a, uint64_t e);  /* :204-225 */
int zo_f_inv(uint64_t p, uint64_t a, uint64_t *out);    /* :157-191 */

/* ---- Multilinear(F): src/poly/multilinear.zig ---- */
int zo_mle_check(uint64_t n, uint32_t *num_vars);                                   /* init :36-54 */
uint64_t zo_mle_sum(uint64_t p, const uint64_t *e, uint64_t n);                    /* sumOverHypercube :188-194 */
int zo_mle_round_poly(uint64_t p, const uint64_t *e, uint64_t n, uint64_t out[2]); /* roundPolynomial :205-232 */
int zo_mle_partial_eval(uint64_t p, const uint64_t *e, uint64_t n, uint64_t r, uint64_t *out /* n/2 */); /* :154-180 */
int zo_mle_eval(uint64_t p, const uint64_t *e, uint64_t n, const uint64_t *point, uint32_t npoint, uint64_t *out); /* :110-144 */
void zo_mle_add(uint64_t p, const uint64_t *a, const uint64_t *b, uint64_t n, uint64_t *out);   /* :235-250 */
void zo_mle_scalar_mul(uint64_t p, const uint64_t *a, uint64_t s, uint64_t n, uint64_t *out);   /* :253-264 */

/* ---- FiatShamirTranscript: src/core/hash.zig:255-324 ---- */
typedef struct { zo_sha3_256 h; } zo_transcript;
void zo_transcript_init(zo_transcript *t);                                   /* :261-276 */
void zo_transcript_append_field(zo_transcript *t, uint64_t value);           /* :279-283 */
void zo_transcript_append_bytes(zo_transcript *t, const void *d, size_t n);  /* :293-295 */
uint64_t zo_transcript_challenge(zo_transcript *t, uint64_t p);              /* :301-316 + digestToFieldElement :228-242 */
void zo_hash_field_element(uint64_t value, uint8_t out[32]);                 /* hashFieldElementSHA3 :135-147 */
void zo_merge_hashes(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]); /* mergeHashesSHA3 :187-195 */

/* ---- Sumcheck: src/proofs/sumcheck_{protocol,prover,verifier}.zig ---- */
uint64_t zo_eval_univariate(uint64_t p, const uint64_t *coeffs, uint32_t n, uint64_t x); /* sumcheck_protocol.zig:113-123 */
/* SumcheckProver.prove (sumcheck_prover.zig:26-91). round_polys: v*2, point: v. */
int zo_sumcheck_prove(uint64_t p, const uint64_t *evals, uint64_t n, uint64_t *round_polys, uint64_t *final_point,
                      uint64_t *final_eval, uint64_t *claimed_sum);
/* proveInteractive (sumcheck_prover.zig:97-144) */
int zo_sumcheck_prove_interactive(uint64_t p, const uint64_t *evals, uint64_t n, const uint64_t *challenges,
                                  uint32_t n_challenges, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval);
/* SumcheckProof.toBytes (sumcheck_protocol.zig:76-109): (2+3v)*8 bytes */
void zo_sumcheck_proof_to_bytes(uint32_t v, const uint64_t *round_polys, const uint64_t *final_point, uint64_t final_eval,
                                uint8_t *out);
/* SumcheckVerifier.verifyRounds (sumcheck_verifier.zig:172-202), generalised to `ncoef` coefficients per round. */
int zo_sumcheck_verify_rounds(uint64_t p, uint32_t v, uint32_t ncoef, const uint64_t *round_polys, uint64_t claimed_sum,
                              int *is_valid, uint64_t *final_claim);

/* ---- OUR EXTENSION in reference conventions: product sumcheck of d MLEs, d in 1..3.
 * Round polynomial in COEFFICIENT form [a0..ad], MSB-first binding, transcript absorbs each coefficient
 * as le64 and then challenge(). d == 1 is bit-identical to zo_sumcheck_prove. No reference behaviour exists for d>1. */
int zo_prod_round_coeffs(uint64_t p, const uint64_t *const *polys, uint32_t d, uint64_t n, uint64_t *out /* d+1 */);
int zo_prodcheck_prove(uint64_t p, const uint64_t *const *polys, uint32_t d, uint64_t n, uint64_t *round_polys /* v*(d+1) */,
                       uint64_t *final_point, uint64_t *final_evals /* d */, uint64_t *claimed_sum);

/* ---- SimpleMerkleTree(F, SHA3Hasher): src/commitments/merkle_tree.zig:273-402 ---- */
uint64_t zo_ceil_pow2(uint64_t n);
/* build :283-318. leaf_hashes: padded*32 bytes (may be NULL), root: 32 bytes. */
int zo_merkle_build(const uint64_t *values, uint64_t n, uint8_t *leaf_hashes, uint8_t root[32], uint32_t *height);
/* open :324-360 — recomputes every level like the reference. siblings: height*32, dirs: height. */
int zo_merkle_open(const uint64_t *values, uint64_t n, const uint8_t *leaf_hashes, uint64_t index, uint8_t *siblings,
                   uint8_t *dirs, uint64_t *value);
/* k openings sharing one recomputation of the levels (test convenience for large trees; same walk as zo_merkle_open) */
int zo_merkle_open_many(const uint64_t *values, uint64_t n, const uint8_t *leaf_hashes, const uint64_t *indices, uint32_t k,
                        uint8_t *siblings, uint8_t *dirs, uint64_t *out_values);
/* verify :362-373 */
int zo_merkle_verify(const uint8_t root[32], uint64_t value, const uint8_t *siblings, const uint8_t *dirs, uint32_t height);

/* ---- CommitmentScheme: src/commitments/polynomial_commit.zig ---- */
uint64_t zo_point_to_index(const uint64_t *point, uint32_t npoint); /* :178-183 */
/* open :86-115 = eval + pointToIndex + tree.open */
int zo_commit_open(uint64_t p, const uint64_t *evals, uint64_t n, const uint8_t *leaf_hashes, const uint64_t *point,
                   uint32_t npoint, uint64_t *value, uint64_t *leaf_index, uint64_t *leaf_value, uint8_t *siblings, uint8_t *dirs);

/* ---- Prover.generateCommitments: src/prover/prover.zig:366-467 (transcript interleave included) ----
 * roots: count*32, points: count*v, values/leaf_indices/leaf_values: count, siblings: count*v*32, dirs: count*v */
int zo_generate_commitments(uint64_t p, zo_transcript *tr, const uint64_t *const *polys, uint32_t count, uint64_t n,
                            uint8_t *roots, uint64_t *points, uint64_t *values, uint64_t *leaf_indices, uint64_t *leaf_values,
                            uint8_t *siblings, uint8_t *dirs);
/* ---- WitnessGenerator.generate on SoA trace columns: src/constraints/witness.zig:29-270. Returns the padded length;
 * out: n_cols * padded */
uint64_t zo_witness_pack(uint64_t p, const uint64_t *cols, uint64_t num_steps, uint32_t n_cols, uint32_t n_hold, uint64_t *out);

/* ---- Prover.prove after the VM (prover.zig:91-110, 156-226) + BinarySerializer.serialize (serialization.zig:70-97) ----
 * cols: the 43 witness columns of the trace (SoA, order of prover.zig:376-390), final_regs: 32 values.
 * compat_buffer != 0 reproduces the reference's under-estimated fixed buffer: error.NoSpaceLeft where `zigz prove` fails. */
uint64_t zo_count_lookups(const uint64_t *opcode_col, uint64_t num_steps);
size_t zo_proof_exact_size(uint64_t num_steps, uint32_t n_init, uint32_t n_out, uint64_t n_lookups);
size_t zo_proof_estimated_size(uint64_t num_steps, uint32_t n_init, uint64_t n_lookups);
int zo_prove_from_trace(uint64_t p, const uint8_t *program, size_t program_len, uint64_t entry_pc, const uint64_t *initial_regs,
                        uint32_t n_init, const uint64_t *cols, uint64_t num_steps, uint64_t final_pc, const uint64_t *final_regs,
                        const uint64_t *outputs, uint32_t n_out, int compat_buffer, uint8_t *out, size_t out_cap, size_t *out_len);
/* Verifier.verify (verifier.zig:49-294) on the serialized proof: *result 0 Accept, 1/2/3 = RejectInvalidSumcheck/Lookup/Commitment */
int zo_verify_proof(uint64_t p, const uint8_t *proof, size_t len, const uint8_t *program, size_t program_len, int *result);

/* ---- Lasso: src/lookups/lasso_prover.zig, table_builder.zig ---- */
/* rows are flattened (inputs || outputs), `arity` u64 per row */
uint64_t zo_lasso_hash_row(uint64_t p, const uint64_t *row, uint32_t arity);          /* hashEntry/hashQuery :208-239 */
void zo_lasso_commit_poly(const uint64_t *evals, uint64_t n, uint8_t out[32]);        /* commitToPolynomial :242-252 */
enum { ZO_TABLE_ADD = 0, ZO_TABLE_XOR = 1, ZO_TABLE_AND = 2 };
/* buildAddTable/XorTable/AndTable :126-213 -> rows (a, b, out), 2^(2*bits) of them */
void zo_build_table(uint64_t p, int op, uint32_t bits, uint64_t *rows);
/* LassoProver.prove :103-173. table length must be a power of two (Multilinear.init). */
int zo_lasso_prove(uint64_t p, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows, uint64_t n_queries,
                   uint32_t arity, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                   uint8_t query_commitment[32], uint8_t table_commitment[32]);
/* proveWithMapping :179-205 (n_inputs + n_outputs = arity; shapes equal by construction of flattened rows) */
int zo_lasso_prove_with_mapping(uint64_t p, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows,
                                uint64_t n_queries, const uint64_t *mapping, uint64_t n_mapping, uint32_t arity,
                                uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                                uint8_t query_commitment[32], uint8_t table_commitment[32]);

/* ---- synthetic inputs shared by tests and bench: x_i = splitmix64(seed + i) mod p ---- */
uint64_t zo_splitmix64(uint64_t x);
void zo_fill_synthetic(uint64_t p, uint64_t seed, uint64_t start, uint64_t n, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
