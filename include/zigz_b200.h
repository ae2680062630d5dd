/* zigz_b200.h — C ABI of the B200 device layer for the zigz proving hot path.
 *
 * The reference (ch4r10t33r/zigz) has no FFI surface: callers use comptime-generic
 * Zig types (src/lib.zig:9-17). Each entry point below replaces the body of the
 * reference function cited next to it; a Zig host binds them with `extern fn`
 * (see INTEGRATION.md). Plain pointers and sizes only; no torch / CUDA types.
 *
 * Conventions
 *   - field elements cross the boundary as the reference's `struct { value: u64 }`
 *     (src/core/field.zig:26-27): canonical BabyBear values in [0, p), 8-byte LE.
 *     On the device they are stored as canonical u32.
 *   - every call returns int32 status: 0 = ok, negative = the reference's Zig error
 *     (names below) or a device error. Nothing falls back to the CPU.
 *   - one context per host thread and GPU; calls on a context are synchronous and
 *     serialized, like the single-threaded reference.
 *   - device objects are opaque 64-bit handles released explicitly.
 */
#ifndef ZIGZ_B200_H
#define ZIGZ_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZB_BABYBEAR_P 2013265921u /* src/core/field_presets.zig:19 */

typedef struct zb_ctx zb_ctx;
typedef uint64_t zb_mle;  /* Multilinear(BabyBear) resident in HBM */
typedef uint64_t zb_tree; /* SimpleMerkleTree(BabyBear, SHA3Hasher) resident in HBM, all levels retained */

enum zb_status {
    ZB_OK = 0,
    ZB_ERR_EMPTY_EVALUATIONS = -1,     /* multilinear.zig:38 */
    ZB_ERR_LENGTH_NOT_POW2 = -2,       /* multilinear.zig:43 */
    ZB_ERR_WRONG_NUM_VARS = -3,        /* multilinear.zig:112 */
    ZB_ERR_NO_VARIABLES = -4,          /* multilinear.zig:156,207; sumcheck_prover.zig:31 */
    ZB_ERR_EMPTY_VALUES = -5,          /* merkle_tree.zig:284 */
    ZB_ERR_INDEX_OUT_OF_BOUNDS = -6,   /* merkle_tree.zig:325 */
    ZB_ERR_POINT_DIM_MISMATCH = -7,    /* polynomial_commit.zig:93 */
    ZB_ERR_NO_QUERIES = -8,            /* lasso_prover.zig:109 */
    ZB_ERR_MAPPING_LEN_MISMATCH = -9,  /* lasso_prover.zig:186 */
    ZB_ERR_INVALID_MAPPING = -10,      /* lasso_prover.zig:192 */
    ZB_ERR_QUERY_TABLE_MISMATCH = -11, /* lasso_prover.zig:199 */
    ZB_ERR_WRONG_NUM_CHALLENGES = -12, /* sumcheck_prover.zig:105 */
    ZB_ERR_DIFFERENT_NUM_VARS = -13,   /* multilinear.zig:237 */
    ZB_ERR_EMPTY_TRACE = -14,          /* prover.zig:145 */
    ZB_ERR_NO_SPACE_LEFT = -15,        /* serialization.zig:72-73: the reference's fixed (under-estimated) buffer */
    ZB_ERR_PROGRAM_HASH_MISMATCH = -16, /* verifier.zig:105 */
    ZB_ERR_INVALID_PROOF = -17,        /* serialization.zig:52-61 deserialize errors */
    ZB_ERR_NOT_CANONICAL = -20,        /* an input element was >= p (reference asserts val < MODULUS, field.zig:43) */
    ZB_ERR_BAD_HANDLE = -21,
    ZB_ERR_BAD_ARGUMENT = -22,
    ZB_ERR_OOM = -100,                 /* error.OutOfMemory */
    ZB_ERR_NO_DEVICE = -200,           /* no CUDA device / driver: there is no CPU fallback */
    ZB_ERR_CUDA = -201,                /* see zb_last_error() */
    ZB_ERR_TIMEOUT = -202,
    ZB_ERR_NCCL = -203                 /* libnccl missing or an NCCL call failed; see zb_last_error() */
};

/* ---- context ---- */
int32_t zb_ctx_create(int32_t device, zb_ctx **out);
/* One context over SEVERAL GPUs of this process (`device_mask` selects 1/2/4/8 GPUs; bit d = device d).
 * No NCCL, no CUDA IPC, no second process: the library enables peer access between the devices, runs one host thread per
 * device, and the kernels exchange their per-round partial sums through each other's memory over NVLink.
 * On such a context zb_mle_upload / _upload_u32 / _synthetic / _constant create SHARDED tables (cyclic layout: GPU = low index
 * bits, so every MSB-first fold pair is local), and zb_mle_free / _len / _download / _sum / _eval, zb_merkle_build / _open /
 * _info / _free, zh_sumcheck_prove*, zh_prodcheck_prove*, zh_commit*, zh_generate_commitments accept them: same results, bit for
 * bit, as on one GPU. Tables need >= world^2 entries. Every other entry point keeps working on plain (single-GPU) handles of
 * the first selected device. A mask with one bit is zb_ctx_create. */
int32_t zb_ctx_create_mask(uint32_t device_mask, zb_ctx **out);
void zb_ctx_destroy(zb_ctx *ctx);
/* number of devices behind ctx (1 for an ordinary context) */
int32_t zb_group_size(zb_ctx *ctx);
/* for host twins: run fn(rank's context, rank, world, user) once per device, concurrently, each on its own host thread;
 * returns the first failing status. zb_group_mle: rank's shard of a sharded table, as a handle of that rank's context. */
typedef int32_t (*zb_rank_fn)(zb_ctx *rank_ctx, int32_t rank, int32_t world, void *user);
int32_t zb_group_run(zb_ctx *ctx, zb_rank_fn fn, void *user);
int32_t zb_group_mle(zb_ctx *ctx, zb_mle h, int32_t rank, zb_mle *out);
int32_t zb_group_mle_set_len(zb_ctx *ctx, zb_mle h, uint64_t n);
const char *zb_last_error(zb_ctx *ctx);
const char *zb_status_name(int32_t status);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t zb_kernel_launches(zb_ctx *ctx);
/* pinned host memory for fast transfers (optional; any host pointer is accepted everywhere) */
int32_t zb_host_alloc(zb_ctx *ctx, size_t bytes, void **out);
int32_t zb_host_free(zb_ctx *ctx, void *p);
int32_t zb_device_info(zb_ctx *ctx, int32_t *sm_count, uint64_t *total_mem, uint64_t *free_mem);
/* measured integer-pipe ceiling of this GPU for the hashing kernels: 32-bit lane-operations per second of independent
 * LOP3 chains, SHF chains, and the Keccak mix (122 LOP3 : 58 SHF), each timed with CUDA events over ~ms-long launches */
int32_t zb_int_pipe_peak(zb_ctx *ctx, double *lop3_per_s, double *shf_per_s, double *keccak_mix_per_s);
/* measured host->device copy rate of this GPU's link (pinned source, one `bytes`-sized cudaMemcpyAsync, best of 3, CUDA
 * events): the denominator of bench.py's pcie_frac */
int32_t zb_h2d_rate(zb_ctx *ctx, size_t bytes, double *bytes_per_s);
/* raw stream handle (cudaStream_t) the context launches on: for CUDA-event timing by a harness */
void *zb_stream(zb_ctx *ctx);
int32_t zb_sync(zb_ctx *ctx);
/* tuning knobs. "tail_log2": tables of <= 2^value entries run all their remaining fold rounds inside ONE persistent
 * kernel that receives the challenges through host-mapped memory (0 disables; default 14, env ZB_TAIL_LOG2). */
/* "comm_reduce" (0/1/2, needs a communicator; 2 needs zb_comm_p2p_attach): zb_prod_round_coeffs / zb_prod_fold_inplace / zb_prod_partial_eval return
 * the coefficients of the WHOLE (sharded) round polynomial: the kernel's partial sums are all-reduced over NCCL in stream
 * order and published once, without a host hop in between. The d final evaluations of the last fold are never summed. */
int32_t zb_set_option(zb_ctx *ctx, const char *key, int64_t value);
int32_t zb_get_option(zb_ctx *ctx, const char *key, int64_t *value);
/* CUDA-event stopwatch on the context's stream (device time between the two calls, in milliseconds) */
int32_t zb_timer_start(zb_ctx *ctx);
int32_t zb_timer_stop(zb_ctx *ctx, float *ms);
/* per-kernel accounting for the roofline: while enabled every launch is bracketed by CUDA events on the launching
 * stream and accumulated by kernel family together with its ALGORITHMIC bytes (DESIGN.md gives the formula per kernel).
 * enable(1) clears the table. */
int32_t zb_profile_enable(zb_ctx *ctx, int32_t on);
uint32_t zb_profile_count(zb_ctx *ctx);
int32_t zb_profile_entry(zb_ctx *ctx, uint32_t i, char *name, uint32_t name_cap, uint64_t *launches, double *total_ms,
                         uint64_t *algorithmic_bytes);

/* ---- Multilinear(F): src/poly/multilinear.zig ---- */
/* init :36-54 — copies `n` evaluations to the device; n must be a power of two, every value < p */
int32_t zb_mle_upload(zb_ctx *ctx, const uint64_t *evals, uint64_t n, zb_mle *out);
/* same, from canonical u32 host values (narrow host representation; optional fast path) */
int32_t zb_mle_upload_u32(zb_ctx *ctx, const uint32_t *evals, uint64_t n, zb_mle *out);
/* zero :57-70 / constant :73-86 */
int32_t zb_mle_constant(zb_ctx *ctx, uint32_t num_vars, uint64_t value, zb_mle *out);
/* synthetic input generated on the device: e[i] = splitmix64(seed + start + i*stride) mod p */
int32_t zb_mle_synthetic(zb_ctx *ctx, uint64_t seed, uint64_t start, uint64_t stride, uint64_t n, zb_mle *out);
int32_t zb_mle_clone(zb_ctx *ctx, zb_mle src, zb_mle *out);
/* deinit :89-91 */
int32_t zb_mle_free(zb_ctx *ctx, zb_mle m);
int32_t zb_mle_len(zb_ctx *ctx, zb_mle m, uint64_t *n, uint32_t *num_vars);
int32_t zb_mle_download(zb_ctx *ctx, zb_mle m, uint64_t *out, uint64_t n);
/* evaluations [offset, offset + n) */
int32_t zb_mle_download_range(zb_ctx *ctx, zb_mle m, uint64_t offset, uint64_t *out, uint64_t n);
/* evaluations [offset, offset + n) in the device representation (canonical u32): a plain D2H copy */
int32_t zb_mle_download_u32(zb_ctx *ctx, zb_mle m, uint64_t offset, uint32_t *out, uint64_t n);
/* a context-owned pinned scratch buffer of at least `bytes` bytes (grown on demand, released with the context) */
int32_t zb_host_scratch(zb_ctx *ctx, size_t bytes, void **out);
/* sumOverHypercube :188-194 */
int32_t zb_mle_sum(zb_ctx *ctx, zb_mle m, uint64_t *out);
/* roundPolynomial :205-232 — returns the two half sums (s0, s1) mod p; the host forms [s0, s1 - s0] */
int32_t zb_mle_round_sums(zb_ctx *ctx, zb_mle m, uint64_t out_s0_s1[2]);
/* partialEval :154-180 — new[i] = (1-r)*e[i] + r*e[i+n/2] (binds the TOP index bit). Returns a NEW polynomial.
 * If next_s0_s1 != NULL the same kernel also produces the next round's half sums (for n/2 == 1 both slots hold
 * the single remaining evaluation and 0). */
int32_t zb_mle_partial_eval(zb_ctx *ctx, zb_mle m, uint64_t r, zb_mle *out, uint64_t next_s0_s1[2]);
/* the same fold, in place: the handle keeps its identity and halves its length (no allocation in the round loop) */
int32_t zb_mle_fold_inplace(zb_ctx *ctx, zb_mle m, uint64_t r, uint64_t next_s0_s1[2]);
/* eval :110-144 — LSB-first: point[k] <-> index bit k. O(N) fold instead of the reference's O(N*v) loop; same value */
int32_t zb_mle_eval(zb_ctx *ctx, zb_mle m, const uint64_t *point, uint32_t npoint, uint64_t *out);
/* eval for `count` polynomials of equal length at one point each (points: count x npoint, row-major) with ONE read-back:
 * Prover.generateCommitments knows all 43 opening points before it needs any value (prover.zig:420-443) */
int32_t zb_mle_eval_batch(zb_ctx *ctx, const zb_mle *polys, uint32_t count, const uint64_t *points, uint32_t npoint, uint64_t *out);
/* add :235-250 / scalarMul :253-264 */
int32_t zb_mle_add(zb_ctx *ctx, zb_mle a, zb_mle b, zb_mle *out);
int32_t zb_mle_scalar_mul(zb_ctx *ctx, zb_mle a, uint64_t scalar, zb_mle *out);

/* ---- product sumcheck rounds (extension in reference conventions) ----
 * g(X) = sum_i prod_k (lo_k[i] + (hi_k[i] - lo_k[i]) X) over the d (1..3) polynomials, MSB-first pairs (i, i+n/2).
 * out_coeffs receives the d+1 COEFFICIENTS [a0..ad] (canonical). d == 1 equals roundPolynomial. */
int32_t zb_prod_round_coeffs(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *out_coeffs);
/* folds all d polynomials in place with r; if next_coeffs != NULL also returns the next round's coefficients.
 * When the folded length is 1, next_coeffs[0..d) receives the d final evaluations instead. */
int32_t zb_prod_fold_inplace(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t r, uint64_t *next_coeffs);
/* the same fold out of place (partialEval semantics for all d polynomials): `out` receives d NEW handles */
int32_t zb_prod_partial_eval(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t r, zb_mle *out, uint64_t *next_coeffs);

/* eq(tau, .) as a table — EXTENSION (no reference behaviour), the weight table of an eq-weighted sumcheck:
 * E[i] = prod_k (bit_k(i) ? tau[k] : 1 - tau[k]), index bit k <-> tau[k] as in Multilinear.eval (multilinear.zig:128-141), hence
 * sum_i E[i] * A[i] == A.eval(tau). One product per entry from two half tables built on the host. */
int32_t zb_mle_eq(zb_ctx *ctx, const uint64_t *tau, uint32_t num_vars, zb_mle *out);

/* ---- two sumcheck rounds per pass over the data ----
 * With the top two index bits as variables (X, Y), G(X, Y) = sum_i prod_k B_k,i(X, Y) (B bilinear through the four quarter
 * elements i, i+n/4, i+n/2, i+3n/4) holds the next TWO round polynomials: g(X) = G(X,0) + G(X,1) and, once the challenge r
 * of that round is known, g'(Y) = G(r, Y). `grid` receives G on P x P with P = {0,1} (d=1), {0,1,inf} (d=2),
 * {0,1,-1,inf} (d=3), row-major grid[ix*|P| + iy], canonical values (summed over the ranks under "comm_reduce").
 * zb_prod_grid: G of the tables as they are (read only).
 * zb_prod_fold_grid: first bind `nfold` (1 or 2) top variables with r[0] (, r[1]) — exactly partialEval applied nfold
 * times — then G of the folded tables. out == NULL folds in place; otherwise d NEW tables are returned and the inputs stay
 * untouched. Needs folded length >= 8 (error.BadArgument otherwise; use the single-round entries for small tables). */
int32_t zb_prod_grid(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *grid);
int32_t zb_prod_fold_grid(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint32_t nfold, const uint64_t *r, zb_mle *out,
                          uint64_t *grid);

/* ---- small product tables leave the device (the last rounds of SumcheckProver.prove, sumcheck_prover.zig:50-77, are latency) ----
 * zb_prod_fold_dump: bind `nfold` (0..2) top variables with r[0] (, r[1]) — partialEval applied nfold times — and return the
 * FOLDED tables to the host: tables[k * m + i], k < d, i < m = n >> nfold, canonical u32; m <= 2^12. The device tables are not
 * modified. The host twin finishes the remaining log2(m) rounds itself: a host round trip per round (~8 us) would cost more
 * than the few hundred field multiplications the round is.
 * zb_prod_collapse: every table becomes the single value values[k] (length 1) — how a consuming prove leaves its tables when
 * the last rounds ran on the host. Stream-ordered, returns without waiting. */
int32_t zb_prod_fold_dump(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint32_t nfold, const uint64_t *r, uint32_t *tables);
int32_t zb_prod_collapse(zb_ctx *ctx, const zb_mle *polys, uint32_t d, const uint64_t *values);

/* ---- several rounds per pass for ONE polynomial (d = 1, SumcheckProver.prove sumcheck_prover.zig:50-77) ----
 * roundPolynomial (multilinear.zig:205-232) is linear in the table, so the 2^k sums over the blocks selected by the top k
 * index bits hold the next k round polynomials (the host folds the 2^k sums with each challenge exactly as partialEval
 * folds the table), and the k partialEval steps (:154-180) that follow collapse into one pass:
 *   new[i] = sum_b w_b e[b m + i],  w_b = prod_j (b_j ? r_j : 1 - r_j),  m = n / 2^k  — the same canonical values.
 * zb_mle_block_sums: sums[b], b < 2^k (1 <= k <= 8, n >= 2^(k+2)); more than 32 blocks are meant for small tables (one CTA
 * per block).
 * zb_mle_fold_multi: binds the top k_fold variables with r[0..k_fold) (r[0] = the top bit, as partialEval would), in
 * place when out == NULL (the handle shrinks to m entries) or into a NEW table, and returns 2^k_next sums over the blocks of
 * the folded table. k_next == log2(m) (allowed up to 12) returns the folded table itself — the host can then finish
 * the remaining <= 12 rounds without another device round trip; k_fold may then be 1..8 (a 2^20-entry prove is two device
 * calls: 256 block sums, then one fold of 8 variables that publishes 2^12 entries). Otherwise k_fold <= 5, k_next <= 5 and
 * m >= 2^(k_next+2). */
int32_t zb_mle_block_sums(zb_ctx *ctx, zb_mle m, uint32_t k, uint64_t *sums);
int32_t zb_mle_fold_multi(zb_ctx *ctx, zb_mle m, uint32_t k_fold, const uint64_t *r, zb_mle *out, uint32_t k_next, uint64_t *sums);
/* replaces the table by the single value `value` (length 1): how a consuming prove leaves its polynomial when the last
 * rounds were finished from a published table */
int32_t zb_mle_collapse(zb_ctx *ctx, zb_mle m, uint64_t value);

/* ---- SimpleMerkleTree(F, SHA3Hasher): src/commitments/merkle_tree.zig:273-402 ---- */
/* build :283-318 for `count` polynomials of equal length in one batch (CommitmentScheme.batchCommit,
 * polynomial_commit.zig:132-157; Prover.generateCommitments, prover.zig:405-410). roots: count*32 bytes. */
int32_t zb_merkle_build(zb_ctx *ctx, const zb_mle *polys, uint32_t count, zb_tree *trees, uint8_t *roots);
/* build from host values of arbitrary length n >= 1 (pads to a power of two with hash(0), :303-306) */
int32_t zb_merkle_build_values(zb_ctx *ctx, const uint64_t *values, uint64_t n, zb_tree *tree, uint8_t root[32]);
int32_t zb_merkle_info(zb_ctx *ctx, zb_tree t, uint64_t *n_values, uint32_t *height, uint8_t root[32]);
/* open :324-360 — siblings: height*32 bytes leaf->root, dirs[l] = (index >> l) & 1, leaf_value = values[index] */
int32_t zb_merkle_open(zb_ctx *ctx, zb_tree t, uint64_t index, uint8_t *siblings, uint8_t *dirs, uint64_t *leaf_value);
/* open for `count` trees of equal shape, one leaf index each, in one launch and one read-back: siblings count x height x 32 bytes,
 * dirs count x height, leaf_values count */
int32_t zb_merkle_open_batch(zb_ctx *ctx, const zb_tree *trees, uint32_t count, const uint64_t *indices, uint8_t *siblings,
                             uint8_t *dirs, uint64_t *leaf_values);
/* leaf_hashes copy-out (tests): padded*32 bytes */
int32_t zb_merkle_leaf_hashes(zb_ctx *ctx, zb_tree t, uint8_t *out, uint64_t n_digests);
int32_t zb_merkle_free(zb_ctx *ctx, zb_tree t);

/* ---- Lasso: src/lookups/lasso_prover.zig:208-239, table_builder.zig:126-213 ---- */
/* hashEntry / hashQuery over flattened rows (inputs || outputs), `arity` u64 each:
 * h = 0; for x in row: h ^= x; h = XXH3_64(seed 0, le64(h)); eval = h % p. Rows beyond n_rows up to n_padded are zero
 * (lasso_prover.zig:140-142). Result is a Multilinear of n_padded (power of two) evaluations. */
int32_t zb_xxh3_rows(zb_ctx *ctx, const uint64_t *rows, uint64_t n_rows, uint32_t arity, uint64_t n_padded, zb_mle *out);
/* The same, streamed: rows go up, are hashed and come back down into `host_mirror` (n_padded canonical u32, from
 * zb_host_mirror) chunk by chunk; after each chunk *avail (release store) is the number of leading evaluations that are
 * final in host_mirror. A second host thread can consume the mirror while the call is still running — this is how
 * LassoProver.prove hides the upload and the sumcheck behind the sequential commitToPolynomial sponge
 * (lasso_prover.zig:160-164, 242-252). On return *avail == n_padded (or the status is an error). */
int32_t zb_xxh3_rows_stream(zb_ctx *ctx, const uint64_t *rows, uint64_t n_rows, uint32_t arity, uint64_t n_padded, zb_mle *out,
                            uint32_t *host_mirror, uint64_t *avail);
/* a second context-owned pinned buffer (independent of zb_host_scratch) of at least `bytes` bytes */
int32_t zb_host_mirror(zb_ctx *ctx, size_t bytes, void **out);
/* buildAddTable (op 0) / buildXorTable (1) / buildAndTable (2) hashed directly on the device:
 * entry index = a * 2^bits + b -> hashEntry((a, b) -> op(a, b)) */
int32_t zb_table_mle(zb_ctx *ctx, int32_t op, uint32_t bits, zb_mle *out);

/* ---- WitnessGenerator.generate: src/constraints/witness.zig:29-270, on a column (SoA) view of the execution trace ----
 * cols: n_cols columns of num_steps raw u64 values each (column-major), in the polynomial order of prover.zig:376-390
 * (pc, x0..x31, opcode, rd, rs1, rs2, funct3, funct7, imm, mem.address, mem.value, mem.is_read => n_cols = 43, n_hold = 33).
 * Every value is reduced mod p (F.init); tables are padded to 2^ceil(log2 num_steps): the first n_hold columns repeat
 * their last value (:80-87, :116-123), the others pad with zero (:174-182, :249-253). out: n_cols new polynomials. */
int32_t zb_witness_pack(zb_ctx *ctx, const uint64_t *cols, uint64_t num_steps, uint32_t n_cols, uint32_t n_hold, zb_mle *out,
                        uint32_t *num_vars);

/* zb_witness_pack and the commit of the n_cols (<= 64) resulting polynomials (zb_merkle_build) as ONE pipeline: while column c
 * crosses PCIe, the GPU packs and leaf-hashes column c-1. Same polynomials, same trees and roots as the two calls in a row. */
int32_t zb_witness_pack_commit(zb_ctx *ctx, const uint64_t *cols, uint64_t num_steps, uint32_t n_cols, uint32_t n_hold, zb_mle *out,
                               uint32_t *num_vars, zb_tree *trees, uint8_t *roots);

/* ---- multi-GPU: one context per process and GPU; NCCL over NVLink/NVSwitch carries the per-round exchange ----
 * The hypercube is sharded CYCLICALLY (rank = low log2(world) index bits) so that every MSB-first pair (i, i + n/2)
 * of partialEval / roundPolynomial is local to one GPU; per round only the d+1 partial coefficients cross GPUs.
 * libnccl is dlopen()ed on first use (`nccl_path` may be NULL: then ZIGZ_NCCL_LIB, then "libnccl.so.2"). */
int32_t zb_comm_unique_id(const char *nccl_path, uint8_t out[128]);
int32_t zb_comm_init(zb_ctx *ctx, const char *nccl_path, const uint8_t unique_id[128], int32_t rank, int32_t world);
int32_t zb_comm_info(zb_ctx *ctx, int32_t *rank, int32_t *world); /* world == 1 when no communicator is attached */
/* exact element-wise sum over all ranks of n (<= 64) u64 values, in place (host memory) */
int32_t zb_comm_allreduce_u64(zb_ctx *ctx, uint64_t *vals, uint32_t n);
/* NVLink peer exchange (optional, on top of the communicator): every rank exports a small exchange buffer over CUDA IPC
 * (zb_comm_p2p_handle, 64 bytes), the handles of all ranks in rank order are attached (zb_comm_p2p_attach, world*64
 * bytes), and "comm_reduce" = 2 then makes the LAST CTA of each round kernel store its partial sums into every peer's
 * buffer, wait for the peers' flags and publish the total — the reduction rides inside the kernel that produced the
 * sums: no NCCL call and no extra launch per round. */
int32_t zb_comm_p2p_handle(zb_ctx *ctx, uint8_t out[64]);
int32_t zb_comm_p2p_attach(zb_ctx *ctx, const uint8_t *handles);
/* all-gather of cyclic shards: out[rank + world*j] = shard_rank[j] (a NEW polynomial of world * n_local entries,
 * identical on every rank) — used to leave the sharded regime once the tables are small */
int32_t zb_comm_allgather_cyclic(zb_ctx *ctx, zb_mle local, zb_mle *out);
/* the same for `count` (1..3) shards of equal length with ONE rendezvous and one stream synchronisation */
int32_t zb_comm_allgather_cyclic_batch(zb_ctx *ctx, const zb_mle *locals, uint32_t count, zb_mle *outs);
int32_t zb_comm_destroy(zb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* ZIGZ_B200_H */
