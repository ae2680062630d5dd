/* zigz_host.h — C exports of the HOST side of the zigz proving hot path.
 *
 * The reference's host language is Zig; no Zig toolchain exists in the build image, so the host bodies that a
 * Zig maintainer would keep (Fiat-Shamir transcript, the sumcheck round loop, proof assembly, the commitment
 * scheme glue, the Lasso driver) are written in C++ ON TOP OF the device C ABI of zigz_b200.h — they call
 * nothing but zb_* entry points — and are exported here with plain-C signatures so that tests (ctypes) and a Zig
 * host can drive them. Function-for-function they mirror the reference API named beside each entry:
 * same argument meaning, same outputs, same error names (enum zb_status).
 *
 * The transcript, the per-round challenge derivation and all proof bytes are produced on the host, exactly as in
 * the reference; the device only ever sees evaluation tables, one challenge per round and digests.
 */
#ifndef ZIGZ_HOST_H
#define ZIGZ_HOST_H
#include "zigz_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- FiatShamirTranscript: src/core/hash.zig:255-324 (host-resident SHA3-256 sponge) ---- */
typedef struct zh_transcript zh_transcript;
zh_transcript *zh_transcript_new(void);                                         /* init :261-276 */
zh_transcript *zh_transcript_clone(const zh_transcript *t);
void zh_transcript_free(zh_transcript *t);
void zh_transcript_append_field(zh_transcript *t, uint64_t value);              /* appendFieldElement :279-283 */
void zh_transcript_append_fields(zh_transcript *t, const uint64_t *v, size_t n); /* appendFieldElements :286-290 */
void zh_transcript_append_bytes(zh_transcript *t, const void *data, size_t n);  /* appendBytes :293-295 */
uint64_t zh_transcript_challenge(zh_transcript *t);                             /* challenge(BabyBear) :301-316 */
void zh_transcript_finalize(zh_transcript *t, uint8_t out[32]);                 /* finalize :319-323 */
/* std.crypto.hash.sha3.Sha3_256 one-shot (hashBytesSHA3, hash.zig:150-160) */
void zh_sha3_256(const void *data, size_t n, uint8_t out[32]);
/* digestToFieldElement(BabyBear) hash.zig:228-242 */
uint64_t zh_digest_to_field(const uint8_t digest[32]);

/* ---- BabyBear = Field(u64, 2013265921): src/core/field.zig (host scalar ops used by the twin) ---- */
uint64_t zh_f_add(uint64_t a, uint64_t b); /* :73-88 */
uint64_t zh_f_sub(uint64_t a, uint64_t b); /* :91-98 */
uint64_t zh_f_mul(uint64_t a, uint64_t b); /* :112-147 */
/* evalUnivariateCoeffs: src/proofs/sumcheck_protocol.zig:113-123 */
uint64_t zh_eval_univariate(const uint64_t *coeffs, uint32_t n, uint64_t x);

/* ---- SumcheckProver(BabyBear): src/proofs/sumcheck_prover.zig ---- */
/* Tuning: tables of at least 2^v entries are proved two rounds per pass over the data (zb_prod_grid / zb_prod_fold_grid:
 * ~10.7 instead of 16 bytes of HBM traffic per element and polynomial, half the host round trips); smaller ones one round per
 * kernel. 0 disables the two-round path; -1 selects the default (env ZB_GRID_MIN_LOG2, else 5 while small product tables finish
 * on the host — zb option "prod_host_tail_log2" > 0 — and 15 otherwise). Returns the previous setting. Process-wide. */
int32_t zh_set_grid_min_log2(int32_t v);
/* prove :26-91. `poly` is left untouched (the reference copies it, :47). round_polys: v*2 coefficients [s0, s1-s0],
 * final_point: v challenges, final_eval: current_poly.evaluations[0], claimed_sum: sumOverHypercube (:40).
 * error.NoVariables when num_vars == 0. */
int32_t zh_sumcheck_prove(zb_ctx *ctx, zb_mle poly, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval,
                          uint64_t *claimed_sum);
/* proveInteractive :97-144 — error.WrongNumberOfChallenges when n_challenges != num_vars */
int32_t zh_sumcheck_prove_interactive(zb_ctx *ctx, zb_mle poly, const uint64_t *challenges, uint32_t n_challenges,
                                      uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval);
/* SumcheckProof.toBytes: src/proofs/sumcheck_protocol.zig:76-109. out: (2 + 3v)*8 bytes; returns the byte count */
size_t zh_sumcheck_proof_to_bytes(uint32_t num_vars, const uint64_t *round_polys, const uint64_t *final_point,
                                  uint64_t final_eval, uint8_t *out);
/* Product sumcheck of d (1..3) polynomials — extension in the reference's conventions:
 * round polynomial in coefficient form [a0..ad], MSB-first binding, the transcript absorbs every coefficient as le64
 * and then challenge(). d == 1 produces exactly zh_sumcheck_prove's output. round_polys: v*(d+1), final_evals: d.
 * The inputs are left untouched. */
int32_t zh_prodcheck_prove(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *round_polys, uint64_t *final_point,
                           uint64_t *final_evals, uint64_t *claimed_sum);
/* Multi-GPU: when a communicator is attached to ctx (zb_comm_init, world = P) the three prove entry points above and
 * below take this rank's CYCLIC shard (local element j = global element rank + P*j) and prove the sum over the whole
 * 2^(v_local + log2 P) hypercube: round_polys / final_point then hold v_local + log2(P) rounds, identical on all ranks. */
/* same, consuming the inputs (folded in place: no extra device memory; handles end with length 1) */
int32_t zh_prodcheck_prove_consume(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *round_polys,
                                   uint64_t *final_point, uint64_t *final_evals, uint64_t *claimed_sum);

/* The host's share of a product prove: the LAST rounds, once the tables are small. tables[k * m + i] (k < d, i < m, m a power of
 * two <= 4096, canonical u32 — what zb_prod_fold_dump returns) are the current tables before round `round`; the function runs
 * rounds round .. round + log2(m) - 1 exactly as the provers above do (roundPolynomial over MSB-first pairs, transcript `tr`
 * or fixed challenges, partialEval; sumcheck_prover.zig:50-77) and writes round_polys[(round + t) * (d + 1) ..], final_point[round + t],
 * final_evals[0..d). `tables` is scratch (folded in place). The provers call this once their tables have <= 2^"prod_host_tail_log2"
 * entries: a device round trip per round (~8 us) costs more than the few hundred multiplications the round is. */
int32_t zh_prodcheck_finish_small(uint32_t d, uint32_t *tables, uint64_t m, uint32_t round, zh_transcript *tr,
                                  const uint64_t *fixed_challenges, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_evals);

/* measurement helper: `reps` zh_sumcheck_prove calls in a row, wall-clock microseconds per prove (no binding overhead) */
int32_t zh_time_sumcheck_prove(zb_ctx *ctx, zb_mle poly, uint32_t reps, double *us_per_prove);

/* EXTENSION, clearly outside the reference (SURVEY.md §8 f4 "real degree-d sumcheck"): eq-weighted product sumcheck
 *     sum_x eq(tau, x) * prod_{k<d} A_k(x),  d = 1 or 2,
 * the form constraint-satisfaction and zero-check arguments use. Proved as the product sumcheck of the d + 1 tables
 * (eq(tau, .), A_0, .., A_{d-1}) in the conventions of zh_prodcheck_prove: round polynomials of degree d + 1 in
 * coefficient form (round_polys: v * (d + 2)), MSB-first binding, same transcript. final_evals: d + 1 values, eq first
 * (== eq(tau, challenges in index-bit order)). tau[k] <-> index bit k as in Multilinear.eval, so for d == 1 the claimed
 * sum is A_0.eval(tau) (multilinear.zig:110-144). The inputs are left untouched. Lasso's grand-product / memory-checking
 * arguments are NOT built: the reference has no such code to be bit-exact with (lasso_prover.zig:147-158 are comments). */
int32_t zh_eqcheck_prove(zb_ctx *ctx, const uint64_t *tau, uint32_t num_vars, const zb_mle *polys, uint32_t d, uint64_t *round_polys,
                         uint64_t *final_point, uint64_t *final_evals, uint64_t *claimed_sum);

/* ---- CommitmentScheme(BabyBear, SHA3Hasher): src/commitments/polynomial_commit.zig ---- */
/* commit :69-83 */
int32_t zh_commit(zb_ctx *ctx, zb_mle poly, zb_tree *tree, uint8_t root[32], uint32_t *num_vars);
/* batchCommit :132-157 */
int32_t zh_batch_commit(zb_ctx *ctx, const zb_mle *polys, uint32_t count, zb_tree *trees, uint8_t *roots);
/* commit of a polynomial sharded over the context's communicator by CONTIGUOUS blocks (rank g holds leaves
 * [g N/P, (g+1) N/P)): each GPU builds its subtree, the P subtree roots are gathered and the top log2(P) levels are
 * hashed on the host. `root` is the SimpleMerkleTree root of the whole polynomial, identical on every rank. */
int32_t zh_commit_sharded(zb_ctx *ctx, zb_mle local_poly, zb_tree *tree, uint8_t local_root[32], uint8_t root[32]);
/* open :86-115 — value = poly.eval(point), leaf_index = pointToIndex(point) (:178-183), Merkle path of that leaf.
 * siblings: num_vars*32 bytes, dirs: num_vars bytes. error.PointDimensionMismatch when npoint != num_vars */
int32_t zh_commit_open(zb_ctx *ctx, zb_mle poly, zb_tree tree, const uint64_t *point, uint32_t npoint, uint64_t *value,
                       uint64_t *leaf_index, uint64_t *leaf_value, uint8_t *siblings, uint8_t *dirs);
/* pointToIndex :178-183 */
uint64_t zh_point_to_index(const uint64_t *point, uint32_t npoint);
/* SimpleMerkleTree.verify: src/commitments/merkle_tree.zig:362-373 (host; O(height) hashes). returns 1 / 0 */
int32_t zh_merkle_verify(const uint8_t root[32], uint64_t value, const uint8_t *siblings, const uint8_t *dirs, uint32_t height);
/* CommitmentScheme.verify :118-129 : opening.value == proof.value is NOT checked there; only the path. returns 1 / 0 */
int32_t zh_commit_verify(const uint8_t root[32], uint64_t leaf_value, const uint8_t *siblings, const uint8_t *dirs,
                         uint32_t height);

/* ---- Prover.generateCommitments: src/prover/prover.zig:366-467 ----
 * commit all `count` (43) witness polynomials, absorb "POLY_COMMITMENTS" + roots, then per polynomial derive its v opening
 * challenges, evaluate, open leaf pointToIndex(point); finally absorb "OPENING_CLAIMS" + values. The transcript is the
 * caller's (it already holds program hash, sumcheck and Lasso traffic). roots: count*32, points: count*v, values /
 * leaf_indices / leaf_values: count, siblings: count*v*32, dirs: count*v. */
int32_t zh_generate_commitments(zb_ctx *ctx, zh_transcript *tr, const zb_mle *polys, uint32_t count, uint8_t *roots,
                                uint64_t *points, uint64_t *values, uint64_t *leaf_indices, uint64_t *leaf_values,
                                uint8_t *siblings, uint8_t *dirs);

/* ---- Prover.prove after the VM + BinarySerializer.serialize: src/prover/prover.zig:91-226, serialization.zig:70-97 ----
 * Everything `zigz prove` does once the execution trace exists: bind program hash / entry pc / initial registers, pack the
 * 43 witness polynomials on the device (zb_witness_pack), placeholder constraint sumcheck and per-lookup Lasso entries
 * with their transcript traffic (prover.zig:229-362), generateCommitments on the device, public I/O, "ZIGZ" v1 bytes.
 * trace_cols: 43 SoA columns x num_steps raw u64 (order of prover.zig:376-390); final_regs: 32 values.
 * compat_buffer != 0 reproduces the reference's under-estimated fixed buffer (error.NoSpaceLeft once num_vars >= 19 or
 * with a few thousand lookup steps, serialization.zig:134-173); 0 serializes any size.
 * out_len always receives the exact size; error.OutOfMemory if out_cap is too small (call once with out = NULL to size). */
int32_t zh_prove_from_trace(zb_ctx *ctx, const uint8_t *program, size_t program_len, uint64_t entry_pc,
                            const uint64_t *initial_regs, uint32_t n_initial_regs, const uint64_t *trace_cols, uint64_t num_steps,
                            uint64_t final_pc, const uint64_t *final_regs, const uint64_t *outputs, uint32_t n_outputs,
                            int32_t compat_buffer, uint8_t *out, size_t out_cap, size_t *out_len);
/* Verifier.verify (src/verifier/verifier.zig:49-294) on serialized proof bytes. *verdict: 0 Accept, 1 RejectInvalidSumcheck,
 * 2 RejectInvalidLookup, 3 RejectInvalidCommitment. error.ProgramHashMismatch / error.InvalidProof as status. Host only. */
int32_t zh_verify_proof(const uint8_t *proof, size_t len, const uint8_t *program, size_t program_len, int32_t *verdict);
/* std.crypto.hash.sha2.Sha256 one-shot (program hash, prover.zig:98-99) */
void zh_sha256(const void *data, size_t n, uint8_t out[32]);

/* Tuning: query lists that pad to at least 2^v evaluations run the query commitment (the sequential SHA3 sponge,
 * lasso_prover.zig:242-252) on a second host thread, fed chunk by chunk through zb_xxh3_rows_stream while the rows are still
 * being uploaded and the sumcheck runs. -1 = never. Default 18 (env ZB_LASSO_PIPELINE_MIN_LOG2). Returns the previous value.
 * Process-wide. Chunk size: zb_set_option(ctx, "lasso_chunk_log2", rows_log2), default 19. */
int32_t zh_set_lasso_pipeline_min_log2(int32_t v);

/* ---- LassoProver(BabyBear): src/lookups/lasso_prover.zig ---- */
/* prove :103-173. Rows are flattened (inputs || outputs), `arity` u64 each (the reference's TableEntry / LookupQuery
 * hold separately allocated slices, table_builder.zig:14-35, lasso_prover.zig:65-86).
 * Outputs: the sumcheck proof of the zero-padded query polynomial (round_polys: v*2, final_point: v), num_vars = v,
 * query_commitment / table_commitment = commitToPolynomial (:242-252, flat SHA3 over le64 evaluations, host).
 * error.NoQueries, error.LengthNotPowerOfTwo (table), error.NoVariables (a single query => v = 0). */
int32_t zh_lasso_prove(zb_ctx *ctx, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows,
                       uint64_t n_queries, uint32_t arity, uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval,
                       uint32_t *num_vars, uint8_t query_commitment[32], uint8_t table_commitment[32]);
/* proveWithMapping :179-205 */
int32_t zh_lasso_prove_with_mapping(zb_ctx *ctx, const uint64_t *table_rows, uint64_t n_table, const uint64_t *query_rows,
                                    uint64_t n_queries, const uint64_t *mapping, uint64_t n_mapping, uint32_t arity,
                                    uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                                    uint8_t query_commitment[32], uint8_t table_commitment[32]);
/* the same proof for a table generated on the device (buildAddTable / buildXorTable / buildAndTable,
 * table_builder.zig:126-213; op 0/1/2, `bits`-wide operands): no host table, no table upload */
int32_t zh_lasso_prove_builtin(zb_ctx *ctx, int32_t op, uint32_t bits, const uint64_t *query_rows, uint64_t n_queries,
                               uint64_t *round_polys, uint64_t *final_point, uint64_t *final_eval, uint32_t *num_vars,
                               uint8_t query_commitment[32], uint8_t table_commitment[32]);
/* n_jobs independent proofs of that kind in one call — e.g. the per-lookup-kind proofs of one execution trace. Same
 * outputs as n_jobs calls of zh_lasso_prove_builtin (per-job arrays; commitments are n_jobs x 32 bytes; statuses[j] is job
 * j's status and the return value the first non-zero one). The query commitments, each one sequential SHA3 sponge
 * (:242-252) and ~98 % of a proof's time, run concurrently on one host thread per proof while the calling thread drives the
 * GPU work of the next proof. Not in the reference, whose prover is single-threaded: an extension for throughput. */
int32_t zh_lasso_prove_builtin_batch(zb_ctx *ctx, uint32_t n_jobs, const int32_t *ops, const uint32_t *bits,
                                     const uint64_t *const *query_rows, const uint64_t *n_queries, uint64_t *const *round_polys,
                                     uint64_t *const *final_points, uint64_t *final_evals, uint32_t *num_vars,
                                     uint8_t *query_commitments, uint8_t *table_commitments, int32_t *statuses);
/* commitToPolynomial :242-252 over host evaluations: SHA3-256 of le64(e[0]) || ... || le64(e[n-1]) (8-byte elements,
 * or canonical 4-byte elements that are absorbed zero-extended) */
void zh_flat_commit(const uint64_t *evals, uint64_t n, uint8_t out[32]);
void zh_flat_commit_u32(const uint32_t *evals, uint64_t n, uint8_t out[32]);
/* commitToPolynomial :242-252 of a device-resident polynomial */
int32_t zh_lasso_commit_poly(zb_ctx *ctx, zb_mle poly, uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif /* ZIGZ_HOST_H */
