// Internal launch interface between the context layer (ctx.cu) and the sm_100a kernels.
// Every launcher enqueues on `st` and returns; results that the host needs come back through the
// Mailbox (mapped pinned memory written by the last CTA of the kernel, then polled by the host).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace zk {

constexpr int MAX_POLYS = 3;    // product sumcheck degree
constexpr int MAX_BATCH = 64;   // trees per batched Merkle launch
constexpr int MAIL_WORDS = 40;  // u64 payload words per mailbox (16-word round grids, 32 block sums, a status word)

// Completion mailbox. `acc`/`ticket` live in device memory; `mail` is mapped pinned host memory.
// mail[0..MAIL_WORDS) payload, mail[MAIL_WORDS] = sequence number written last (release, system scope).
// Peer exchange over NVLink (multi-GPU): every rank owns one exchange buffer that all peers have mapped (CUDA IPC).
// Two alternating sets (seq parity); per set XCHG_MAX_RANKS payload rows of XCHG_ROW words + one flag word per source rank.
constexpr int XCHG_MAX_RANKS = 16;
constexpr int XCHG_ROW = 16;
constexpr int XCHG_SET_WORDS = XCHG_MAX_RANKS * XCHG_ROW + XCHG_MAX_RANKS;
// Behind the two payload sets the same buffer carries the staging area of the in-kernel all-gather that ends the sharded regime
// (launch_gather_xchg): per set XCHG_MAX_RANKS arrival words, then per set MAX_POLYS shards of up to 2^XCHG_GATHER_MAX_LOG2 u32.
constexpr int XCHG_GATHER_MAX_LOG2 = 16;
constexpr size_t XCHG_GATHER_FLAGS_OFF = 2 * (size_t)XCHG_SET_WORDS;                    // in u64 words
constexpr size_t XCHG_GATHER_DATA_OFF = XCHG_GATHER_FLAGS_OFF + 2 * XCHG_MAX_RANKS;     // in u64 words
constexpr size_t XCHG_GATHER_SET_ELEMS = (size_t)3 << XCHG_GATHER_MAX_LOG2;            // u32 per set (3 = MAX_POLYS)
constexpr size_t XCHG_BUFFER_BYTES = XCHG_GATHER_DATA_OFF * 8 + 2 * XCHG_GATHER_SET_ELEMS * 4;
struct XchgView {
    unsigned long long *peer[XCHG_MAX_RANKS]; // peer[q] = rank q's exchange buffer as seen from this GPU (peer[rank] = own)
    int rank, world;
    long long patience; // SM cycles a rank waits for its peers' rows before it reports a starved exchange
    unsigned long long *stats; // device: {cycles spent waiting for peers, exchanged rounds}, accumulated (nullptr: off)
};

struct Mailbox {
    unsigned long long *acc;    // device: MAIL_WORDS u64 accumulators (zero between launches)
    unsigned int *ticket;       // device: CTA arrival counter (zero between launches)
    unsigned long long *mail;   // device alias of the mapped host mailbox
    unsigned long long seq;     // sequence value this launch must publish
    const XchgView *xchg;       // non-null: sum the payload over all ranks through peer memory before publishing
    unsigned long long xseq;    // exchange round number: counts exchanged rounds only, identical on every rank
    // tagged payload: every payload word is written as (low 32 bits of seq) << 32 | value (all payload values are canonical
    // field elements < 2^31), so each word validates itself and NO system-scope fence is needed between the payload and the
    // sequence number — the host waits until every word it expects carries the tag. Saves ~1.5 us per host round trip
    // (profiles/r02_sweep7_*.txt). false: plain words, fence, then the sequence number (payloads that are not 32-bit values:
    // digests, published tables; and mailboxes that are reduced over ranks first).
    bool tagged;
};
__host__ __device__ inline unsigned long long mail_tagged(unsigned long long seq, unsigned long long v) {
    return ((seq & 0xffffffffull) << 32) | (v & 0xffffffffull);
}

struct PolySet {
    const uint32_t *src[MAX_POLYS];
    uint32_t *dst[MAX_POLYS];
};

// Round sums of the current polynomials over MSB-first pairs (i, i + n/2), n >= 2.
//   D=1: payload {s0, s1}            D=2: {g(0), g(1), g(inf)}      D=3: {g(0), g(1), g(-1), g(inf)}
// all canonical mod p.
void launch_round_sums(int d, const PolySet &ps, uint64_t n, const Mailbox &mb, int sm_count, cudaStream_t st);

// Where a fold kernel gets its challenge from. chal == nullptr: the launch parameter r (immediate).
// Otherwise the kernel was PRE-LAUNCHED behind the previous round's kernel, before the host knew the challenge: the first
// CTA to arrive (atomic claim) polls the host-mapped word *chal until it holds (tag << 32 | r), republishes it in the
// device word *bcast for the other CTAs, and the fold starts — the launch latency of the round is off the critical path.
// Tag 0xFFFFFFFF aborts (every CTA leaves without touching the tables); ~0.2 s without a challenge does the same.
struct ChalSrc {
    const unsigned long long *chal;
    unsigned long long *bcast;
    unsigned int *claim;
    unsigned int tag;
};

// Fold every polynomial with challenge r (new[i] = e[i] + r (e[i + n/2] - e[i])) writing n/2 values to dst
// (dst may equal src: in place), and, fused, the round sums of the folded polynomials (same payload as above).
// If n == 2 the payload is the d final evaluations instead.
void launch_fold_sums(int d, const PolySet &ps, uint64_t n, uint32_t r, const Mailbox &mb, int sm_count, cudaStream_t st,
                      const ChalSrc *cs = nullptr);
// true when launch_fold_sums(n) runs the vectorised kernel (the only one that supports a polled challenge)
inline bool fold_sums_is_vector(uint64_t n) { return n >= 16 && ((n / 4) % 4) == 0; }

// Persistent single-CTA kernel that performs ALL remaining fold rounds of d polynomials of length n (in place,
// ps.dst == ps.src). Round k (k = 0, 1, ...) waits until the host-mapped word *chal holds (chal_seq0 + k) << 32 | r_k,
// folds with r_k and publishes the same payload as launch_fold_sums with sequence number mb.seq + k.
// A tag of 0xFFFFFFFF aborts (arrays stay consistent with the rounds completed); *status = 1 on a 2 s starvation exit.
void launch_tail_rounds(int d, const PolySet &ps, uint64_t n, const Mailbox &mb, const unsigned long long *chal,
                        unsigned int chal_seq0, unsigned int *status, cudaStream_t st);

// Small product tables leave the device: bind nfold (0..2) top variables with r1 (, r2) and write the m = n >> nfold folded values
// of each of the d tables, canonical u32, to dump[k * m + i] (host-mapped); then fence + mb.seq (no payload words). One CTA;
// m <= 2^PROD_DUMP_MAX_LOG2. The device tables are not modified.
constexpr int PROD_DUMP_MAX_LOG2 = 12;
void launch_fold_dump(int d, int nfold, const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, uint32_t *dump, const Mailbox &mb,
                      cudaStream_t st);
// dst[k][0] = vals[k], k < d
void launch_fill_heads(const PolySet &ps, int d, const uint32_t *vals, cudaStream_t st);

// multi-GPU: src[0..n) (NCCL-summed canonical payload words) -> mailbox payload mod p + sequence number
void launch_publish_reduced(const unsigned long long *src, int n, unsigned long long *mail, unsigned long long seq, cudaStream_t st);
// multi-GPU all-gather of `count` cyclic shards (n_local <= 2^XCHG_GATHER_MAX_LOG2 entries each) in ONE kernel over peer memory:
// every rank copies its shards into its own exchange buffer, raises its arrival word in every peer's buffer, waits for the
// peers' words and pulls their shards into the global order outs.dst[k][q + world * j] = shard_q,k[j]. gseq: gather round
// number, identical on all ranks, alternating between the two staging sets. A peer that never arrives sets the mailbox status
// word (mail[MAIL_WORDS - 1] = 1), which the next wait on the mailbox reports. No host rendezvous, no stream synchronisation.
void launch_gather_xchg(const XchgView *xv, int world, const PolySet &shards, const PolySet &outs, int count, uint64_t n_local,
                        unsigned long long gseq, unsigned int *ticket, unsigned long long *mail, cudaStream_t st);
// multi-GPU: all-gathered cyclic shards [rank][j] -> global order out[rank + world * j]
void launch_interleave(const uint32_t *gathered, uint32_t *out, uint64_t n_local, uint32_t world, cudaStream_t st);

// single-process multi-GPU: peer-memory transposes between the cyclic (sumcheck) and contiguous (host order / Merkle) layouts
struct PeerSrc {
    const uint32_t *p[XCHG_MAX_RANKS];
};
struct PeerDst {
    uint32_t *p[XCHG_MAX_RANKS];
};
// out[j * world + q] = src.p[q][j], j < n_local (reads the peers)
void launch_interleave_peers(const PeerSrc &src, uint32_t *out, uint64_t n_local, uint32_t world, int sm_count, cudaStream_t st);
// dst.p[q][j] = src[j * world + q], j < n_out (writes into the peers)
void launch_deal_peers(const uint32_t *src, const PeerDst &dst, uint64_t n_out, uint32_t world, int sm_count, cudaStream_t st);

// Two rounds per pass. With the top two index bits of the (possibly folded) tables as variables (X, Y), the bivariate
//   G(X, Y) = sum_i prod_k B_k,i(X, Y),   B bilinear through the four quarter elements,
// holds BOTH next round polynomials: g(X) = G(X, 0) + G(X, 1) and, once r is known, g'(Y) = G(r, Y). One pass over the
// data therefore serves two sumcheck rounds: `nfold` (0..2) variables are bound first with r1 (and r2), the folded
// tables (length m = n / 2^nfold) are written to dst (in place allowed), and G of the folded tables is published on the
// grid P x P, P = {0,1} (D=1), {0,1,inf} (D=2), {0,1,-1,inf} (D=3): payload[ix * |P| + iy], canonical.
// Needs m >= 8 and m % 8 == 0. HBM traffic per prove drops from 16 d N to ~10.7 d N bytes, host round trips halve.
void launch_fold_grid(int d, int nfold, const PolySet &ps, uint64_t m, uint32_t r1, uint32_t r2, const Mailbox &mb, int sm_count,
                      cudaStream_t st);
inline bool fold_grid_ok(uint64_t m) { return m >= 8 && (m % 8) == 0; }

// ---- d = 1 (the reference's own prover): several rounds per pass through LINEARITY ----
// roundPolynomial of one multilinear table is linear in the table: s0/s1 of round t are sums of the 2^k block sums
// S[b] = sum of e over block b (top k index bits = b) folded with the challenges so far. So 2^k block sums serve the next k
// rounds on the host, and the device then binds those k variables in ONE pass as a 2^k-term dot product with the eq weights
// w_b = prod_j (b_j ? r_j : 1 - r_j) — exactly partialEval applied k times (exact field arithmetic) — emitting the next
// block sums on the way: 2^28 entries take 5 passes (8.3 B per element) instead of 28 rounds (13.3-16 B per element).
constexpr int LIN_MAX_K = 5;          // variables per pass (32 block sums / 32-term dot product)
constexpr int LIN_DUMP_MAX_LOG2 = 12; // a folded table of <= 2^12 entries is published whole (block length 1)
constexpr int LIN_WIDE_MAX_K = 8;     // small tables: up to 8 variables per pass (launch_block_sums_wide / launch_foldk_wide)
struct FoldWeights {
    uint32_t w[1 << LIN_MAX_K]; // Montgomery form (w R mod p), index = the K top index bits (first-bound variable = MSB)
};
// payload[b] = sum of src over block b of nb = 2^k (1..5) equal blocks, canonical; n % (4 nb) == 0
void launch_block_sums(const uint32_t *src, uint64_t n, int k, const Mailbox &mb, int sm_count, cudaStream_t st);
// dst[i] = sum_t w[t] src[t m + i] (m = n >> k, i < m; dst may equal src). k_next >= 0 with m >> k_next >= 4: payload =
// 2^k_next block sums of dst. dump != nullptr (needs m <= 2^LIN_DUMP_MAX_LOG2, m % 4 == 0): the folded values themselves go to
// dump[0..m) as canonical u32 (host-mapped, 16-byte aligned) instead, and the payload is empty.
void launch_foldk_sums(const uint32_t *src, uint32_t *dst, uint64_t n, int k, const FoldWeights &w, int k_next,
                       unsigned long long *dump, const Mailbox &mb, int sm_count, cudaStream_t st);

// The same two steps for SMALL tables (latency-bound: fewer, wider passes), k up to LIN_WIDE_MAX_K. Results go to `words`
// (host-mapped u64) as self-validating words (low 32 bits of seq) << 32 | value — no ticket, no fence, no sequence word: the host
// waits until every word carries the tag.
//   block sums: words[b] = sum of src over block b of 2^k equal blocks (n % (4 * 2^k) == 0), one CTA per block
//   fold:       dst[i] = sum_b w[b] src[b m + i] (m = n >> k <= 2^LIN_DUMP_MAX_LOG2, m % 4 == 0; dst may equal src) and words[i] = dst[i]
struct WideWeights {
    uint32_t w[1 << LIN_WIDE_MAX_K]; // Montgomery form, index as in FoldWeights
};
void launch_block_sums_wide(const uint32_t *src, uint64_t n, int k, unsigned long long *words, unsigned long long seq, cudaStream_t st);
void launch_foldk_wide(const uint32_t *src, uint32_t *dst, uint64_t n, int k, const WideWeights &w, unsigned long long *words,
                       unsigned long long seq, cudaStream_t st);

// Plain sum of all n evaluations: payload {sum mod p}
void launch_sum(const uint32_t *src, uint64_t n, const Mailbox &mb, int sm_count, cudaStream_t st);

// LSB-first evaluation (Multilinear.eval): one stage folds up to 12 variables: out[j] = fold of src[4096 j ..)
// `point` are canonical challenges for the variables this stage folds (lowest first).
struct EvalPoint {
    uint32_t r[12];
    uint32_t rp[12];
};
void launch_eval_stage(const uint32_t *src, uint64_t n, int nvars, const EvalPoint &pt, uint32_t *out, const Mailbox *mb,
                       int sm_count, cudaStream_t st);

// The remaining nv + nv2 (nv <= 12, nv2 <= 8) variables of n = 2^(nv + nv2) values in one launch: tiles of 2^nv per CTA into
// out[0 .. 2^nv2), then the last CTA folds those with pt2 and publishes the value.
void launch_eval_finish(const uint32_t *src, uint64_t n, int nv, const EvalPoint &pt, uint32_t *out, int nv2, const EvalPoint &pt2,
                        const Mailbox &mb, int sm_count, cudaStream_t st);

// Batched Multilinear.eval (count polynomials of n entries, one point each): one stage folds tiles of 2^nv (nv <= 12) entries of
// every polynomial: out[p][tile]. srcs: device array of count table pointers, or nullptr when the tables are the rows of
// `src_rows` (row stride n). pts: device array [count][2][v] = (challenge, Shoup companion) per variable; this stage uses the
// variables var0 .. var0 + nv. publish != nullptr (n == 2^nv): the count results also go to publish[p] (host-mapped u64) and the
// last CTA raises mb.seq.
void launch_eval_stage_batch(const uint32_t *const *srcs, const uint32_t *src_rows, uint64_t n, uint32_t count, int nv,
                             const uint32_t *pts, uint32_t v, uint32_t var0, uint32_t *out, unsigned long long *publish,
                             const Mailbox &mb, cudaStream_t st);
// the 10-variable warp stage for large tables (n a multiple of 1024), batched the same way
void launch_eval_warp10_batch(const uint32_t *const *srcs, uint64_t n, uint32_t count, const uint32_t *pts, uint32_t v, uint32_t var0,
                              uint32_t *out, int sm_count, cudaStream_t st);

// Large stage: folds the 10 low variables of `src` (n a multiple of 1024): out[j] = fold of src[1024 j ..). pt.r[0..10).
void launch_eval_warp10(const uint32_t *src, uint64_t n, const EvalPoint &pt, uint32_t *out, int sm_count, cudaStream_t st);

// element-wise helpers
void launch_narrow_u64(const uint64_t *src, uint32_t *dst, uint64_t n, unsigned int *err_flag, cudaStream_t st);
void launch_check_u32(const uint32_t *src, uint64_t n, unsigned int *err_flag, cudaStream_t st);
void launch_widen_u32(const uint32_t *src, uint64_t *dst, uint64_t n, cudaStream_t st);
void launch_fill(uint32_t *dst, uint64_t n, uint32_t value, cudaStream_t st);
void launch_synthetic(uint32_t *dst, uint64_t n, uint64_t seed, uint64_t start, uint64_t stride, cudaStream_t st);
void launch_add(const uint32_t *a, const uint32_t *b, uint32_t *out, uint64_t n, cudaStream_t st);
// out[i] = lo[i & (2^s - 1)] * hi_mont[i >> s] * 2^-32 mod p
void launch_eq_table(const uint32_t *lo, const uint32_t *hi_mont, int s, uint64_t n, uint32_t *out, cudaStream_t st);
void launch_scalar_mul(const uint32_t *a, uint32_t s, uint32_t *out, uint64_t n, cudaStream_t st);

// Witness packing (witness.zig:29-270): see k_witness_pack. n_cols <= 64.
void launch_witness_pack(const uint64_t *cols, uint64_t chunk, uint64_t step0, uint64_t num_steps, uint64_t padded,
                         uint32_t n_cols, uint32_t n_hold, const uint32_t *last_vals, uint32_t *const *out_cols, cudaStream_t st);

// Lasso row hashing (hashEntry / hashQuery). rows: n_rows * arity canonical u32 on the device; out: n_padded u32.
void launch_xxh3_rows(const uint32_t *rows, uint64_t n_rows, uint32_t arity, uint64_t n_padded, uint32_t *out, cudaStream_t st);
void launch_table_mle(int op, uint32_t bits, uint32_t *out, cudaStream_t st);

// Merkle. Tree storage = all levels concatenated: level l starts at digest offset level_offset(padded, l).
struct MerkleBatch {
    const uint32_t *values[MAX_BATCH]; // per tree
    uint64_t n_values[MAX_BATCH];      // real (unpadded) number of values per tree
    uint8_t *tree[MAX_BATCH];          // per tree storage base
    uint32_t count;
};
inline uint64_t merkle_level_offset(uint64_t padded, uint32_t level) { // in digests
    return 2 * padded - (2 * padded >> level);
}
// integer-pipe microbenchmark (int_peak.cu): mode 0 LOP3, 1 SHF, 2 Keccak-like mix; 64 ops per thread and iteration
void launch_int_peak(int mode, uint32_t *out, int iters, int ctas, cudaStream_t st);
void keccak_init_constants(); // once per context, before the first hashing launch
void launch_merkle_leaves(const MerkleBatch &b, uint64_t padded, cudaStream_t st);
// hashes level `level` -> `level + 1` for every tree of the batch
void launch_merkle_level(const MerkleBatch &b, uint64_t padded, uint32_t level, cudaStream_t st);
// finishes all levels from `level` (width <= 1024) to the root inside one CTA per tree
void launch_merkle_top(const MerkleBatch &b, uint64_t padded, uint32_t level, cudaStream_t st);
constexpr uint64_t MERKLE_TOP_WIDTH = 1024;
// copies the `height` (<= 64) sibling digests of leaf `index` (leaf -> root) to `out` (host-mapped bulk area), then
// publishes mb.seq
void launch_merkle_path(const uint8_t *tree, uint64_t padded, uint32_t height, uint64_t index, uint8_t *out, const Mailbox &mb,
                        cudaStream_t st);
// `count` (<= 65535) openings in one launch: paths to out[t][height][32], leaf values to vals[t]; trees / values / idx are
// device arrays of `count` entries
void launch_merkle_path_batch(const uint8_t *const *trees, const uint32_t *const *values, const uint64_t *idx, uint32_t count,
                              uint64_t padded, uint32_t height, uint8_t *out, uint32_t *vals, cudaStream_t st);
// copies the root of every tree of the batch to `out` (host-mapped bulk area), then publishes mb.seq
void launch_merkle_roots(const MerkleBatch &b, uint64_t padded, uint8_t *out, const Mailbox &mb, cudaStream_t st);

} // namespace zk
