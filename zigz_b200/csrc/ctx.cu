// Context layer: implements the C ABI of include/zigz_b200.h on top of the kernels (kernels.h).
// Owns the stream, a caching device allocator (no cudaMalloc inside the sumcheck round loop), the
// host-mapped completion mailbox, and the handle tables. There is no CPU fallback anywhere:
// without a CUDA device zb_ctx_create fails with ZB_ERR_NO_DEVICE.
#include "../../include/zigz_b200.h"
#include "bb.cuh"
#include "hostpack.hpp"
#include "kernels.h"
#include "sha3_host.hpp"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <immintrin.h>
#include <mutex>
#include <cstdlib>
#include <dlfcn.h>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace zk;

namespace {

struct DevBuf {
    struct zb_ctx *ctx;
    void *ptr;
    size_t bytes;
    DevBuf(zb_ctx *c, void *p, size_t b) : ctx(c), ptr(p), bytes(b) {}
    ~DevBuf();
};
using BufRef = std::shared_ptr<DevBuf>;

struct Mle {
    BufRef buf;
    uint64_t n;
    uint32_t *d() const { return (uint32_t *)buf->ptr; }
};

struct Tree {
    BufRef values;      // u32 values: the tree's own copy (merkle_tree.zig:291)
    uint64_t n_values;  // unpadded
    uint64_t padded;
    uint32_t height;
    BufRef store;       // all levels
    uint8_t root[32];
};

constexpr size_t BULK_OFFSET = 512;          // bytes into the mapped mailbox page (payload words + sequence number come first)
constexpr size_t BULK_BYTES = 64 * 32 * 2;   // 64 digests for paths / roots (x2 slack)
constexpr size_t DUMP_OFFSET = BULK_OFFSET + BULK_BYTES; // published tables of zb_mle_fold_multi (u64 per value)
constexpr size_t DUMP_BYTES_LIN = sizeof(unsigned long long) << LIN_DUMP_MAX_LOG2;
constexpr size_t DUMP_BYTES_PROD = (MAX_POLYS * sizeof(uint32_t)) << PROD_DUMP_MAX_LOG2; // zb_prod_fold_dump: d tables of u32
constexpr size_t DUMP_BYTES = DUMP_BYTES_LIN > DUMP_BYTES_PROD ? DUMP_BYTES_LIN : DUMP_BYTES_PROD;
static_assert((MAIL_WORDS + 1) * sizeof(unsigned long long) <= BULK_OFFSET, "mailbox payload overlaps the bulk area");
constexpr size_t STAGE_ELEMS = 32ull << 20;  // upload staging chunk (u64 elements)

} // namespace

struct zb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    // mailbox
    unsigned long long *h_mail = nullptr; // mapped pinned
    unsigned long long *d_mail = nullptr; // device alias
    unsigned long long *d_acc = nullptr;
    unsigned int *d_ticket = nullptr;
    unsigned int *d_err = nullptr;
    unsigned long long seq = 0;
    uint64_t launches = 0;
    uint64_t h2d_bytes = 0; // bytes handed to host->device copies by the upload paths (bench.py: pcie_frac)
    cudaStream_t copy_stream = nullptr; // second stream for copies that overlap kernels (zb_witness_pack_commit), made on demand
    cudaEvent_t pipe_up[3] = {nullptr, nullptr, nullptr}, pipe_used[3] = {nullptr, nullptr, nullptr};
    // allocator cache
    std::multimap<size_t, void *> free_blocks;
    size_t cached_bytes = 0;
    // handles
    std::unordered_map<uint64_t, Mle> mles;
    std::unordered_map<uint64_t, Tree> trees;
    uint64_t next_handle = 1;
    std::string last_error;
    // stopwatch + per-kernel accounting
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool profiling = false;
    struct ProfEntry {
        std::string name;
        uint64_t launches = 0;
        double ms = 0;
        uint64_t bytes = 0;
    };
    struct ProfPending {
        cudaEvent_t a, b;
        uint32_t entry;
    };
    std::vector<ProfEntry> prof;
    std::vector<ProfPending> prof_pending;
    std::vector<cudaEvent_t> event_pool;
    // persistent tail session (launch_tail_rounds): host -> device challenge word, mapped pinned
    unsigned long long *h_chal = nullptr, *d_chal = nullptr;
    unsigned int chal_seq = 0;
    unsigned int *d_tail_status = nullptr;
    struct Tail {
        bool active = false;
        uint32_t d = 0;
        zb_mle h[3] = {0, 0, 0};
        uint64_t n = 0;
    } tail;
    bool tail_test_starve = false;
    // pre-launched fold kernel (ChalSrc): queued behind the current round's kernel, waiting for its challenge
    struct Pre {
        bool active = false;
        uint32_t d = 0;
        zb_mle h[3] = {0, 0, 0};
        uint64_t n = 0;
        unsigned int tag = 0;
        Mailbox mb{};
        bool red = false;
    } pre;
    // default off: measured neutral on one GPU (8.3533 vs 8.3537 ms per 2^30 prove, profiles/r01_prelaunch.txt) — the
    // PCIe poll of the challenge costs what the launch it replaces cost. Kept (tested) as an option: ZB_PRELAUNCH=1.
    bool prelaunch = false;
    int starved = 0; // polling kernels that left without their challenge
    unsigned long long *d_bcast = nullptr; // device word the polling CTA republishes the challenge in
    unsigned int *d_claim = nullptr;
    int tail_log2 = 14; // tables of <= 2^tail_log2 entries finish inside the persistent kernel (0 = never)
    // d = 1 provers: several rounds per pass through linearity (zb_mle_block_sums / zb_mle_fold_multi); once the folded
    // table has <= 2^host_tail_log2 entries it is published whole and the host finishes the (latency-bound) last rounds
    bool linear_d1 = true;
    int host_tail_log2 = LIN_DUMP_MAX_LOG2;
    // product provers (d <= 3): tables of <= 2^this entries are handed to the host twin, which finishes the rounds (0: never)
    int prod_host_tail_log2 = 10;
    int linear_k = LIN_MAX_K; // variables bound per pass
    void *scratch = nullptr; // zb_host_scratch
    size_t scratch_bytes = 0;
    void *mirror = nullptr; // zb_host_mirror
    size_t mirror_bytes = 0;
    int lasso_chunk_log2 = 19; // rows per chunk of zb_xxh3_rows_stream
    // host-narrowing upload path
    std::unique_ptr<zigz::HostPool> pool;
    uint32_t *pack_buf[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pack_done[3] = {nullptr, nullptr, nullptr};
    int pack_next = 0;
    // multi-GPU
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;
    unsigned long long *d_comm = nullptr; // 64 u64 exchange buffer
    unsigned long long *h_comm = nullptr; // pinned twin
    int comm_reduce = 0; // 0: per-rank payloads; 1: NCCL all-reduce in stream order; 2: NVLink peer exchange inside the kernel
    // NVLink peer exchange (CUDA IPC): own buffer, the peers' mappings, and the device-resident view the kernels read
    unsigned long long *d_xchg = nullptr;
    void *xchg_peer[16] = {nullptr};
    XchgView *d_xchg_view = nullptr;
    unsigned long long gather_seq = 0; // rounds of the in-kernel all-gather (identical on every rank)
    unsigned long long *d_xchg_stats = nullptr; // {wait cycles, exchanged rounds}
    // single-process multi-GPU (zb_ctx_create_mask): the contexts of one process that form a communicator without NCCL
    // or CUDA IPC share a LocalComm (host barrier + pointer exchange for the peer-memory collectives); the FRONT context
    // (rank 0, the one handed to the caller) additionally owns the group: children, worker threads, sharded handles
    std::shared_ptr<struct LocalComm> local;
    struct Group *group = nullptr;
    unsigned long long xchg_seq = 0;

    Mailbox mailbox() {
        Mailbox m;
        m.acc = d_acc;
        m.ticket = d_ticket;
        m.mail = d_mail;
        if (((++seq) & 0xffffffffull) == 0) ++seq; // tag 0 is what a cleared mailbox word carries
        m.seq = seq;
        m.xchg = nullptr;
        m.xseq = 0;
        m.tagged = true; // payload words validate themselves (kernels.h); kernels that publish other things ignore it
        return m;
    }
    uint8_t *h_bulk() { return (uint8_t *)h_mail + BULK_OFFSET; }
    uint8_t *d_bulk() { return (uint8_t *)d_mail + BULK_OFFSET; }
};

// ---- single-process multi-GPU ----
struct LocalComm {
    int world = 1;
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    const void *ptr[XCHG_MAX_RANKS] = {nullptr};
    const void *ptrs[MAX_POLYS][XCHG_MAX_RANKS] = {{nullptr}}; // batched all-gather: [table][rank]
    // sense-reversing barrier between the ranks' host threads (every rank calls it the same number of times)
    void barrier() {
        const int s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == world - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s ^ 1, std::memory_order_release);
        } else {
            while (sense.load(std::memory_order_acquire) == s) _mm_pause();
        }
    }
};
struct GMle { // one table sharded CYCLICALLY over the group's GPUs (rank = low index bits)
    zb_mle child[XCHG_MAX_RANKS];
    uint64_t n; // total length
};
struct GTree { // one Merkle tree sharded by contiguous subtree
    zb_tree child[XCHG_MAX_RANKS];
    uint64_t n_values;
    uint32_t height;
    std::vector<uint8_t> top; // the top log2(world) levels: (2 world - 1) digests, level l (width world >> l) at offset 2 world - (2 world >> l)
    uint8_t root[32];
};
struct Group {
    int world = 1;
    zb_ctx *child[XCHG_MAX_RANKS] = {nullptr}; // child[0] is the front context itself
    std::vector<std::thread> workers;          // ranks 1 .. world-1; rank 0 runs on the calling thread
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int32_t(int)> job;
    uint64_t epoch = 0;
    int pending = 0;
    bool stop = false;
    int32_t status[XCHG_MAX_RANKS] = {0};
    std::unordered_map<uint64_t, GMle> mles;
    std::unordered_map<uint64_t, GTree> trees;
    uint64_t next_handle = 1;
};
constexpr uint64_t GROUP_HANDLE_BIT = 1ull << 63;
static thread_local bool t_in_group_job = false; // inside a per-rank job the front context is just rank 0's context
static inline bool group_front(const zb_ctx *c) { return c->group != nullptr && !t_in_group_job; }
static inline bool is_group_handle(uint64_t h) { return (h & GROUP_HANDLE_BIT) != 0; }
// runs job(rank) once per rank, concurrently (rank 0 on the calling thread); returns the first failing status
static int32_t group_run(zb_ctx *front, const std::function<int32_t(int)> &job);

namespace {

int32_t cuda_fail(zb_ctx *c, cudaError_t e, const char *what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    if (c) c->last_error = buf;
    return e == cudaErrorMemoryAllocation ? ZB_ERR_OOM : ZB_ERR_CUDA;
}
#define CK(call)                                               \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call); \
    } while (0)

size_t round_size(size_t b) {
    if (b < 512) return 512;
    // round to 1/16 of the enclosing power of two so the cache reuses blocks across rounds
    size_t p = 1;
    while (p < b) p <<= 1;
    size_t step = p / 16;
    return (b + step - 1) / step * step;
}

void release_cache(zb_ctx *c) {
    for (auto &kv : c->free_blocks) cudaFree(kv.second);
    c->free_blocks.clear();
    c->cached_bytes = 0;
}

int32_t dev_alloc(zb_ctx *ctx, size_t bytes, BufRef *out) {
    size_t want = round_size(bytes);
    auto it = ctx->free_blocks.lower_bound(want);
    if (it != ctx->free_blocks.end() && it->first <= want + want / 4) {
        void *p = it->second;
        size_t sz = it->first;
        ctx->free_blocks.erase(it);
        ctx->cached_bytes -= sz;
        *out = std::make_shared<DevBuf>(ctx, p, sz);
        return ZB_OK;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        release_cache(ctx);
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError(); // the failed allocation must not surface again at the next launch check
        return cuda_fail(ctx, e, "cudaMalloc");
    }
    *out = std::make_shared<DevBuf>(ctx, p, want);
    return ZB_OK;
}

} // namespace

DevBuf::~DevBuf() {
    // all work on a context is stream-ordered and every API call drains the stream before returning,
    // so a block can be recycled as soon as its last owner drops it
    if (ctx && ptr) {
        ctx->free_blocks.emplace(bytes, ptr);
        ctx->cached_bytes += bytes;
    }
}

namespace {

void rearm(zb_ctx *c);

// Has the launch with sequence number `seq` published? tagged_words > 0: the kernel wrote that many self-validating payload
// words (Mailbox::tagged) and no fence — every one of them must carry the tag; 0: payload, fence, then the sequence number.
inline bool mail_ready(const zb_ctx *ctx, unsigned long long seq, int tagged_words) {
    volatile const unsigned long long *m = ctx->h_mail;
    if (tagged_words <= 0) return m[MAIL_WORDS] == seq;
    const unsigned long long tag = seq & 0xffffffffull;
    for (int k = 0; k < tagged_words; k++)
        if ((m[k] >> 32) != tag) return false;
    return true;
}
// payload word k of a tagged (or plain 32-bit) mailbox
inline uint64_t mail_word(const zb_ctx *ctx, int k) { return ctx->h_mail[k] & 0xffffffffull; }

// Wait for the mailbox of the most recent launch (published by its last CTA).
int32_t wait_mail_raw(zb_ctx *ctx, unsigned long long seq, int tagged_words);
int32_t wait_mail(zb_ctx *ctx, unsigned long long seq, int tagged_words = 0) {
    const int32_t rc = wait_mail_raw(ctx, seq, tagged_words);
    if (rc) rearm(ctx);
    return rc;
}
int32_t wait_mail_raw(zb_ctx *ctx, unsigned long long seq, int tagged_words) {
    auto t0 = std::chrono::steady_clock::now();
    uint64_t spins = 0;
    while (!mail_ready(ctx, seq, tagged_words)) {
        if ((++spins & 0xFFFF) == 0) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) return cuda_fail(ctx, q, "kernel");
            if (q == cudaSuccess && !mail_ready(ctx, seq, tagged_words)) {
                // stream drained but flag missing: give the PCIe write a moment, then report
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) {
                    ctx->last_error = "mailbox sequence not published";
                    return ZB_ERR_TIMEOUT;
                }
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
                ctx->last_error = "timeout waiting for kernel";
                return ZB_ERR_TIMEOUT;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (ctx->comm_reduce == 2 && ctx->d_xchg_view && ctx->h_mail[MAIL_WORDS - 1] == 1ull) {
        ctx->h_mail[MAIL_WORDS - 1] = 0;
        ctx->last_error = "peer exchange: a rank never arrived";
        return ZB_ERR_TIMEOUT;
    }
    return ZB_OK;
}

// Wait until `count` self-validating words (written by many CTAs, in any order) all carry the tag of `seq`.
int32_t wait_tagged_words(zb_ctx *ctx, const volatile unsigned long long *w, unsigned long long seq, uint64_t count) {
    const unsigned long long tag = seq & 0xffffffffull;
    auto t0 = std::chrono::steady_clock::now();
    uint64_t spins = 0, k = 0;
    while (k < count) {
        if ((w[k] >> 32) == tag) {
            k++;
            continue;
        }
        if ((++spins & 0xFFFF) == 0) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            int32_t rc = ZB_OK;
            if (q != cudaSuccess && q != cudaErrorNotReady) rc = cuda_fail(ctx, q, "kernel");
            else if (q == cudaSuccess && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2) && (w[k] >> 32) != tag) {
                ctx->last_error = "published words never arrived";
                rc = ZB_ERR_TIMEOUT;
            } else if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
                ctx->last_error = "timeout waiting for kernel";
                rc = ZB_ERR_TIMEOUT;
            }
            if (rc) {
                rearm(ctx);
                return rc;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return ZB_OK;
}

int32_t comm_publish(zb_ctx *ctx, unsigned long long seq, int nwords); // defined with the NCCL layer

// Mailbox for a kernel whose payload must be summed over the ranks: the kernel writes into the device exchange
// buffer, comm_publish() then all-reduces it in stream order and publishes the reduced words to the host mailbox.
inline bool reduce_on_device(const zb_ctx *c) { return c->comm_reduce && c->world > 1 && (c->nccl_comm || c->local); }
inline bool reduce_p2p(const zb_ctx *c) { return c->comm_reduce == 2 && c->d_xchg_view; }
inline Mailbox round_mailbox(zb_ctx *c, bool reduce) {
    Mailbox m = c->mailbox();
    if (reduce) {
        m.tagged = false; // plain words: they are summed over the ranks before they reach the host
        if (reduce_p2p(c)) {
            m.tagged = true; // ... by the kernel itself, which then publishes the totals as tagged words
            m.xchg = c->d_xchg_view;
            m.xseq = ++c->xchg_seq;
        } else {
            m.mail = c->d_comm; // the kernel writes the exchange buffer, comm_publish() does the rest
        }
    }
    return m;
}

int32_t check_launch(zb_ctx *ctx, const char *what) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, what);
    return ZB_OK;
}
#define LAUNCHED(what)                       \
    do {                                     \
        int32_t _r = check_launch(ctx, what); \
        if (_r) return _r;                   \
    } while (0)

cudaEvent_t take_event(zb_ctx *c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// drains finished event pairs into the table (all of them when `block`)
void prof_drain(zb_ctx *c, bool block) {
    size_t keep = 0;
    for (size_t i = 0; i < c->prof_pending.size(); i++) {
        auto &pp = c->prof_pending[i];
        if (block) cudaEventSynchronize(pp.b);
        if (block || cudaEventQuery(pp.b) == cudaSuccess) {
            float ms = 0;
            cudaEventElapsedTime(&ms, pp.a, pp.b);
            c->prof[pp.entry].ms += ms;
            c->event_pool.push_back(pp.a);
            c->event_pool.push_back(pp.b);
        } else {
            c->prof_pending[keep++] = pp;
        }
    }
    c->prof_pending.resize(keep);
}

// brackets one kernel launch with events on the launching stream while profiling is on
struct ProfScope {
    zb_ctx *c;
    cudaEvent_t a = nullptr;
    uint32_t entry = 0;
    ProfScope(zb_ctx *ctx, const char *name, uint64_t bytes) : c(ctx) {
        if (!c->profiling) return;
        for (entry = 0; entry < c->prof.size(); entry++)
            if (c->prof[entry].name == name) break;
        if (entry == c->prof.size()) {
            c->prof.emplace_back();
            c->prof.back().name = name;
        }
        c->prof[entry].launches++;
        c->prof[entry].bytes += bytes;
        a = take_event(c);
        cudaEventRecord(a, c->stream);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEvent_t b = take_event(c);
        cudaEventRecord(b, c->stream);
        c->prof_pending.push_back({a, b, entry});
        if (c->prof_pending.size() > 256) prof_drain(c, false);
    }
};

// Ends a running persistent tail kernel (abort tag) so that other work can use the stream. The tables stay
// consistent: every round it completed was written back in place and the handle lengths were updated per round.
// Every public entry runs this first (directly or through tail_quiesce): the context's device becomes the calling
// thread's current device, so two contexts on different GPUs can be driven from one thread, and a host (or torch) that
// switched devices in between cannot send our launches and allocations to the wrong GPU.
inline void ensure_device(zb_ctx *c) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device) cudaSetDevice(c->device);
}

// After a timeout or a CUDA error the kernel that should have re-armed the accumulators may never have published:
// drain the stream and zero them so that the next launch starts clean instead of adding to stale partial sums.
void rearm(zb_ctx *c) {
    cudaStreamSynchronize(c->stream);
    cudaGetLastError();
    cudaMemsetAsync(c->d_acc, 0, MAIL_WORDS * sizeof(unsigned long long), c->stream);
    cudaMemsetAsync(c->d_ticket, 0, 2 * sizeof(unsigned int), c->stream); // ticket + err flag
    cudaStreamSynchronize(c->stream);
}

void tail_quiesce(zb_ctx *c) {
    ensure_device(c);
    if (!c->tail.active && !c->pre.active) return;
    __atomic_store_n(c->h_chal, (unsigned long long)0xFFFFFFFFu << 32, __ATOMIC_RELEASE);
    cudaStreamSynchronize(c->stream);
    __atomic_store_n(c->h_chal, 0ull, __ATOMIC_RELEASE); // tag 0 is never issued: the next polling kernel must not see the abort
    c->tail.active = false;
    c->pre.active = false; // a pre-launched kernel leaves without touching the tables or the exchange counters ...
    if (c->pre.red && c->pre.mb.xchg) c->xchg_seq--; // ... so its exchange round number is handed back
    c->pre.red = false;
}

Mle *get_mle(zb_ctx *ctx, zb_mle h) {
    auto it = ctx->mles.find(h);
    return it == ctx->mles.end() ? nullptr : &it->second;
}
Tree *get_tree(zb_ctx *ctx, zb_tree h) {
    auto it = ctx->trees.find(h);
    return it == ctx->trees.end() ? nullptr : &it->second;
}

int32_t new_mle(zb_ctx *ctx, uint64_t n, zb_mle *out, Mle **m) {
    BufRef b;
    int32_t rc = dev_alloc(ctx, n * sizeof(uint32_t), &b);
    if (rc) return rc;
    uint64_t h = ctx->next_handle++;
    ctx->mles[h] = Mle{b, n};
    *out = h;
    if (m) *m = &ctx->mles[h];
    return ZB_OK;
}

int32_t check_pow2(uint64_t n) {
    if (n == 0) return ZB_ERR_EMPTY_EVALUATIONS; // multilinear.zig:38
    if (n & (n - 1)) return ZB_ERR_LENGTH_NOT_POW2; // multilinear.zig:43
    return ZB_OK;
}

int32_t read_err_flag(zb_ctx *ctx) {
    unsigned int flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->d_err, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
        return ZB_ERR_NOT_CANONICAL;
    }
    return ZB_OK;
}

// Upload of host u64 field elements into a device u32 table.
//  host-narrow mode (default when >= 4 host threads are available): worker threads pack u64 -> u32 into pinned
//    staging buffers (canonical check included) while the previous chunk's H2D copy is in flight: 4 bytes per
//    element cross PCIe instead of 8;
//  direct mode: the u64 data is copied as is and narrowed by a kernel (k_narrow).
// elements per staging buffer (default 2^23: 32 MiB narrowed; ZB_PACK_CHUNK_LOG2 = 16..24 for experiments)
static const uint64_t PACK_CHUNK = [] {
    const char *e = getenv("ZB_PACK_CHUNK_LOG2");
    const int x = e && *e ? atoi(e) : 23;
    return 1ull << (x < 16 ? 16 : (x > 24 ? 24 : x));
}();
constexpr int PACK_BUFS = 3;

int upload_threads(zb_ctx *ctx) {
    if (const char *e = getenv("ZB_UPLOAD_THREADS")) return atoi(e) < 1 ? 1 : atoi(e);
    int hw = (int)std::thread::hardware_concurrency();
    int t = hw / (ctx->world > 0 ? ctx->world : 1);
    return t > 32 ? 32 : (t < 1 ? 1 : t); // packing is memory-bound on the pool's hosts well before 32 threads
}

int32_t ensure_pack(zb_ctx *ctx) {
    if (!ctx->pool) ctx->pool.reset(new zigz::HostPool(upload_threads(ctx)));
    if (!ctx->pack_buf[0]) {
        for (int b = 0; b < PACK_BUFS; b++) {
            CK(cudaHostAlloc((void **)&ctx->pack_buf[b], PACK_CHUNK * sizeof(uint32_t), cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&ctx->pack_done[b], cudaEventDisableTiming));
        }
    }
    return ZB_OK;
}

// Plain H2D copy of `bytes` bytes that does not depend on the source being page-locked: pageable sources are copied by
// the pool threads into the rotating pinned staging buffers while the previous piece is in flight (the driver's own
// pageable path manages ~11 GB/s on these hosts), pinned sources go straight to the DMA engine.
int32_t staged_h2d(zb_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st = nullptr) {
    if (!st) st = ctx->stream;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, h_src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned || bytes < (1u << 20)) {
        CK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        ctx->h2d_bytes += bytes;
        return ZB_OK;
    }
    int32_t rc = ensure_pack(ctx);
    if (rc) return rc;
    const size_t piece = PACK_CHUNK * sizeof(uint32_t);
    const int T = ctx->pool->size();
    for (size_t off = 0; off < bytes; off += piece) {
        const size_t m = bytes - off < piece ? bytes - off : piece;
        const int buf = ctx->pack_next;
        ctx->pack_next = (ctx->pack_next + 1) % PACK_BUFS;
        CK(cudaEventSynchronize(ctx->pack_done[buf]));
        char *stage = (char *)ctx->pack_buf[buf];
        const char *src = (const char *)h_src + off;
        ctx->pool->run([&](int tid) {
            const size_t per = ((m + T - 1) / T + 63) & ~(size_t)63;
            const size_t lo = per * tid < m ? per * tid : m, hi = lo + per < m ? lo + per : m;
            if (hi > lo) memcpy(stage + lo, src + lo, hi - lo);
        });
        CK(cudaMemcpyAsync((char *)d_dst + off, stage, m, cudaMemcpyHostToDevice, st));
        ctx->h2d_bytes += m;
        CK(cudaEventRecord(ctx->pack_done[buf], st));
    }
    return ZB_OK;
}

int32_t upload_narrow_host(zb_ctx *ctx, const uint64_t *host, uint64_t n, uint32_t *dst, bool pinned) {
    {
        int32_t rc0 = ensure_pack(ctx);
        if (rc0) return rc0;
    }
    // Hybrid: the pack threads are the bottleneck (~70-80 GB/s of source on 16 cores) while PCIe still has headroom at
    // 4 bytes per element, so some chunks of a PINNED source skip the CPU: they are copied as 8-byte words by the DMA
    // engine while the threads pack the following chunks, and narrowed by a kernel (k_narrow). Which chunks go raw is
    // decided on the fly from a model of the copy queue: a chunk goes raw when the copies already queued would finish
    // before the threads could pack it (the DMA engine would idle otherwise). ZB_UPLOAD_RAW_EVERY=k (k > 0) forces the
    // fixed pattern "every k-th chunk raw", 0 turns raw chunks off.
    static const int raw_every_cfg = [] {
        const char *e = getenv("ZB_UPLOAD_RAW_EVERY");
        return e && *e ? atoi(e) : -1;
    }();
    static const double pcie_bps = [] {
        const char *e = getenv("ZB_UPLOAD_PCIE_GBS"); // H2D rate the queue model assumes
        return (e && *e ? atof(e) : 53.0) * 1e9;
    }();
    static const int raw_cap_cfg = [] {
        const char *e = getenv("ZB_UPLOAD_RAW_CAP"); // adaptive rule: at most one raw chunk in this many (0: no cap)
        return e && *e ? atoi(e) : -1;
    }();
    // default: capped with a full complement of packing threads (>= 16: the threads nearly keep the link busy on their own, and
    // raw chunks then mostly take DRAM bandwidth away from them); uncapped with fewer threads per GPU (several ranks share the
    // host: 12 threads per rank on the 2-GPU box — the link has headroom, the cores do not: 426 ms per step uncapped against
    // 540 ms with one raw chunk in six, profiles/r02_upload_policy_boxes.txt)
    const int raw_cap = raw_cap_cfg >= 0 ? raw_cap_cfg : (ctx->pool->size() >= 16 ? 6 : 0);
    uint64_t raw_count = 0;
    const int raw_every = pinned ? raw_every_cfg : 0;
    BufRef raw_stage;
    if (raw_every != 0 && n > PACK_CHUNK * 4) {
        int32_t rc = dev_alloc(ctx, PACK_CHUNK * sizeof(uint64_t), &raw_stage);
        if (rc) return rc;
    }
    using clk = std::chrono::steady_clock;
    const auto t_start = clk::now();
    auto now_s = [&] { return std::chrono::duration<double>(clk::now() - t_start).count(); };
    double dma_free_at = 0.0; // model: when the copies queued so far will have drained
    double pack_s = 0.0;      // duration of the last full-chunk pack
    auto queued = [&](uint64_t bytes) {
        const double t = now_s();
        dma_free_at = (dma_free_at > t ? dma_free_at : t) + (double)bytes / pcie_bps;
    };
    const int T = ctx->pool->size();
    std::vector<char> bad(T, 0);
    int buf = ctx->pack_next;
    uint64_t chunk_idx = 0;
    for (uint64_t off = 0; off < n; off += PACK_CHUNK, chunk_idx++) {
        const uint64_t m = n - off < PACK_CHUNK ? n - off : PACK_CHUNK;
        bool raw = false;
        if (raw_stage && raw_every > 0) raw = (chunk_idx % raw_every) == (uint64_t)(raw_every - 1);
        else if (raw_stage) {
            // ... and at most one chunk in raw_cap goes raw: a raw chunk takes twice the link time and twice the DMA reads of
            // host DRAM of a packed one, and on hosts where DRAM rather than the cores limits the packing the uncapped rule
            // sent 40 % raw and LOST 15 % against packing everything (profiles/r02_upload_policy_boxes.txt)
            const double est = pack_s > 0 ? pack_s : (double)m * 8 / (4.0e9 * T);
            raw = dma_free_at - now_s() < est && (raw_cap <= 0 || (raw_count + 1) * (uint64_t)raw_cap <= chunk_idx + 1);
        }
        if (raw) raw_count++;
        if (raw) {
            CK(cudaMemcpyAsync(raw_stage->ptr, host + off, m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d_bytes += m * sizeof(uint64_t);
            queued(m * sizeof(uint64_t));
            {
                ProfScope _ps(ctx, "narrow_u64", m * 12);
                launch_narrow_u64((const uint64_t *)raw_stage->ptr, dst + off, m, ctx->d_err, ctx->stream);
            }
            LAUNCHED("narrow");
            continue;
        }
        CK(cudaEventSynchronize(ctx->pack_done[buf])); // the copy that last used this staging buffer has finished
        uint32_t *stage = ctx->pack_buf[buf];
        const uint64_t *src = host + off;
        const double t_pack = now_s();
        ctx->pool->run([&](int tid) {
            const uint64_t per = ((m + T - 1) / T + 15) & ~15ull;
            const uint64_t lo = per * tid < m ? per * tid : m, hi = lo + per < m ? lo + per : m;
            if (hi > lo && zigz::narrow_u64_to_u32(src + lo, stage + lo, hi - lo, bb::P)) bad[tid] = 1;
        });
        if (m == PACK_CHUNK) pack_s = now_s() - t_pack;
        CK(cudaMemcpyAsync(dst + off, stage, m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d_bytes += m * sizeof(uint32_t);
        queued(m * sizeof(uint32_t));
        CK(cudaEventRecord(ctx->pack_done[buf], ctx->stream));
        buf = (buf + 1) % PACK_BUFS;
    }
    ctx->pack_next = buf;
    int32_t rc = raw_stage ? read_err_flag(ctx) : ZB_OK; // also drains the stream
    CK(cudaStreamSynchronize(ctx->stream));
    if (rc) return rc;
    for (char b : bad)
        if (b) return ZB_ERR_NOT_CANONICAL;
    return ZB_OK;
}

int32_t upload_narrow(zb_ctx *ctx, const uint64_t *host, uint64_t n, uint32_t *dst) {
    static const int mode = [] {
        const char *e = getenv("ZB_UPLOAD_MODE"); // "host" | "device"
        return e && !strcmp(e, "device") ? 0 : (e && !strcmp(e, "host") ? 1 : -1);
    }();
    if (mode == 1) {
        cudaPointerAttributes attr{};
        const bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        return upload_narrow_host(ctx, host, n, dst, pinned);
    }
    if (mode == -1 && n >= (1u << 16)) {
        // measured on the pool's hosts (tools/upload_bench.py, profiles/r01_upload_modes.txt): packing on the host beats
        // the direct 8-byte copy from PINNED memory only with >= ~12 threads per GPU (70.7 vs 54.5 GB/s at 16), but beats
        // the driver's staged copy from PAGEABLE memory already with 3-4 threads (33 vs 11 GB/s)
        cudaPointerAttributes attr{};
        const bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const int t = upload_threads(ctx);
        if ((pinned && t >= 12) || (!pinned && t >= 3)) return upload_narrow_host(ctx, host, n, dst, pinned);
    }
    uint64_t chunk = n < STAGE_ELEMS ? n : STAGE_ELEMS;
    BufRef stage;
    int32_t rc = dev_alloc(ctx, chunk * sizeof(uint64_t), &stage);
    if (rc) return rc;
    for (uint64_t off = 0; off < n; off += chunk) {
        uint64_t m = n - off < chunk ? n - off : chunk;
        CK(cudaMemcpyAsync(stage->ptr, host + off, m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d_bytes += m * sizeof(uint64_t);
        {
            ProfScope _ps(ctx, "narrow_u64", m * 12);
            launch_narrow_u64((const uint64_t *)stage->ptr, dst + off, m, ctx->d_err, ctx->stream);
        }
        LAUNCHED("narrow");
    }
    return read_err_flag(ctx);
}

} // namespace

// multi-device context: operations on sharded tables / trees (defined at the end of this file)
template <typename T>
static int32_t zg_upload(zb_ctx *front, const T *evals, uint64_t n, zb_mle *out);
static int32_t group_make_mle(zb_ctx *front, uint64_t n_total, zb_mle *out, const std::function<int32_t(int, zb_ctx *, zb_mle *)> &make);
static int32_t zg_mle_free(zb_ctx *front, zb_mle h);
static int32_t zg_mle_download(zb_ctx *front, zb_mle h, uint64_t offset, uint64_t *out, uint64_t n);
static int32_t zg_mle_eval(zb_ctx *front, zb_mle h, const uint64_t *point, uint32_t npoint, uint64_t *out);
static int32_t zg_mle_sum(zb_ctx *front, zb_mle h, uint64_t *out);
static int32_t zg_merkle_build(zb_ctx *front, const zb_mle *polys, uint32_t count, zb_tree *trees, uint8_t *roots);
static int32_t zg_merkle_open(zb_ctx *front, zb_tree h, uint64_t index, uint8_t *siblings, uint8_t *dirs, uint64_t *leaf_value);
static int32_t zg_merkle_free(zb_ctx *front, zb_tree h);
static GMle *get_gmle(zb_ctx *front, zb_mle h);
static GTree *get_gtree(zb_ctx *front, zb_tree h);

extern "C" {

const char *zb_status_name(int32_t s) {
    switch (s) {
    case ZB_OK: return "ok";
    case ZB_ERR_EMPTY_EVALUATIONS: return "EmptyEvaluations";
    case ZB_ERR_LENGTH_NOT_POW2: return "LengthNotPowerOfTwo";
    case ZB_ERR_WRONG_NUM_VARS: return "WrongNumberOfVariables";
    case ZB_ERR_NO_VARIABLES: return "NoVariables";
    case ZB_ERR_EMPTY_VALUES: return "EmptyValues";
    case ZB_ERR_INDEX_OUT_OF_BOUNDS: return "IndexOutOfBounds";
    case ZB_ERR_POINT_DIM_MISMATCH: return "PointDimensionMismatch";
    case ZB_ERR_NO_QUERIES: return "NoQueries";
    case ZB_ERR_MAPPING_LEN_MISMATCH: return "MappingLengthMismatch";
    case ZB_ERR_INVALID_MAPPING: return "InvalidMapping";
    case ZB_ERR_QUERY_TABLE_MISMATCH: return "QueryTableMismatch";
    case ZB_ERR_WRONG_NUM_CHALLENGES: return "WrongNumberOfChallenges";
    case ZB_ERR_DIFFERENT_NUM_VARS: return "DifferentNumberOfVariables";
    case ZB_ERR_EMPTY_TRACE: return "EmptyTrace";
    case ZB_ERR_NO_SPACE_LEFT: return "NoSpaceLeft";
    case ZB_ERR_PROGRAM_HASH_MISMATCH: return "ProgramHashMismatch";
    case ZB_ERR_INVALID_PROOF: return "InvalidProof";
    case ZB_ERR_NOT_CANONICAL: return "NotCanonical";
    case ZB_ERR_BAD_HANDLE: return "BadHandle";
    case ZB_ERR_BAD_ARGUMENT: return "BadArgument";
    case ZB_ERR_OOM: return "OutOfMemory";
    case ZB_ERR_NO_DEVICE: return "NoCudaDevice";
    case ZB_ERR_CUDA: return "CudaError";
    case ZB_ERR_TIMEOUT: return "Timeout";
    case ZB_ERR_NCCL: return "NcclError";
    default: return "Unknown";
    }
}

int32_t zb_ctx_create(int32_t device, zb_ctx **out) {
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return ZB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return ZB_ERR_BAD_ARGUMENT;
    zb_ctx *ctx = new zb_ctx();
    ctx->device = device;
    auto fail = [&](cudaError_t err, const char *what) {
        fprintf(stderr, "zigz_b200: %s: %s\n", what, cudaGetErrorString(err));
        cudaGetLastError();
        zb_ctx_destroy(ctx); // releases whatever had been created (every member starts out null)
        return (int32_t)ZB_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    size_t mail_bytes = DUMP_OFFSET + DUMP_BYTES;
    if ((e = cudaHostAlloc((void **)&ctx->h_mail, mail_bytes, cudaHostAllocMapped)) != cudaSuccess) return fail(e, "cudaHostAlloc");
    memset(ctx->h_mail, 0, mail_bytes);
    if ((e = cudaHostGetDevicePointer((void **)&ctx->d_mail, ctx->h_mail, 0)) != cudaSuccess) return fail(e, "cudaHostGetDevicePointer");
    if ((e = cudaMalloc(&ctx->d_acc, MAIL_WORDS * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&ctx->d_ticket, 2 * sizeof(unsigned int))) != cudaSuccess) return fail(e, "cudaMalloc");
    ctx->d_err = ctx->d_ticket + 1;
    if ((e = cudaHostAlloc((void **)&ctx->h_chal, 64, cudaHostAllocMapped)) != cudaSuccess) return fail(e, "cudaHostAlloc");
    memset(ctx->h_chal, 0, 64);
    if ((e = cudaHostGetDevicePointer((void **)&ctx->d_chal, ctx->h_chal, 0)) != cudaSuccess) return fail(e, "cudaHostGetDevicePointer");
    if ((e = cudaMalloc(&ctx->d_tail_status, sizeof(unsigned int))) != cudaSuccess) return fail(e, "cudaMalloc");
    cudaMemsetAsync(ctx->d_tail_status, 0, sizeof(unsigned int), ctx->stream);
    if (const char *t = getenv("ZB_TAIL_LOG2")) ctx->tail_log2 = atoi(t);
    if (const char *t = getenv("ZB_PRELAUNCH")) ctx->prelaunch = atoi(t) != 0;
    if (const char *t = getenv("ZB_LINEAR_D1")) ctx->linear_d1 = atoi(t) != 0;
    if (const char *t = getenv("ZB_LINEAR_K")) {
        const int x = atoi(t);
        ctx->linear_k = x < 1 ? 1 : (x > LIN_MAX_K ? LIN_MAX_K : x);
    }
    if (const char *t = getenv("ZB_HOST_TAIL_LOG2")) {
        const int x = atoi(t);
        ctx->host_tail_log2 = x < 2 ? 2 : (x > LIN_DUMP_MAX_LOG2 ? LIN_DUMP_MAX_LOG2 : x);
    }
    if (const char *t = getenv("ZB_PROD_HOST_TAIL_LOG2")) {
        const int x = atoi(t);
        ctx->prod_host_tail_log2 = x < 0 ? 0 : (x > PROD_DUMP_MAX_LOG2 ? PROD_DUMP_MAX_LOG2 : x);
    }
    if ((e = cudaMalloc(&ctx->d_bcast, 2 * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc");
    ctx->d_claim = (unsigned int *)(ctx->d_bcast + 1);
    cudaMemsetAsync(ctx->d_bcast, 0, 2 * sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(ctx->d_acc, 0, MAIL_WORDS * sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(ctx->d_ticket, 0, 2 * sizeof(unsigned int), ctx->stream);
    keccak_init_constants();
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "init");
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e, "init constants");
    *out = ctx;
    return ZB_OK;
}

void zb_ctx_destroy(zb_ctx *ctx) {
    if (!ctx) return;
    if (ctx->group) { // multi-device front: stop the rank threads, then the other ranks' contexts, then this one
        Group *g = ctx->group;
        ctx->group = nullptr;
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->stop = true;
        }
        g->cv_go.notify_all();
        for (auto &t : g->workers) t.join();
        for (int r = 0; r < g->world; r++) cudaStreamSynchronize(g->child[r]->stream);
        for (int r = g->world - 1; r >= 1; r--) zb_ctx_destroy(g->child[r]);
        delete g;
    }
    cudaSetDevice(ctx->device);
    if (ctx->h_chal && ctx->stream) tail_quiesce(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->mles.clear();
    ctx->trees.clear();
    release_cache(ctx);
    zb_comm_destroy(ctx);
    prof_drain(ctx, true);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    cudaFree(ctx->d_acc);
    cudaFree(ctx->d_ticket);
    cudaFree(ctx->d_tail_status);
    cudaFree(ctx->d_bcast);
    if (ctx->scratch) cudaFreeHost(ctx->scratch);
    if (ctx->mirror) cudaFreeHost(ctx->mirror);
    for (int b = 0; b < 3; b++) {
        if (ctx->pack_buf[b]) cudaFreeHost(ctx->pack_buf[b]);
        if (ctx->pack_done[b]) cudaEventDestroy(ctx->pack_done[b]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int b = 0; b < 3; b++) {
        if (ctx->pipe_up[b]) cudaEventDestroy(ctx->pipe_up[b]);
        if (ctx->pipe_used[b]) cudaEventDestroy(ctx->pipe_used[b]);
    }
    if (ctx->h_chal) cudaFreeHost(ctx->h_chal);
    if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
}

const char *zb_last_error(zb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t zb_kernel_launches(zb_ctx *ctx) {
    if (!ctx) return 0;
    uint64_t n = ctx->launches;
    if (group_front(ctx))
        for (int r = 1; r < ctx->group->world; r++) n += ctx->group->child[r]->launches;
    return n;
}
void *zb_stream(zb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int32_t zb_sync(zb_ctx *ctx) {
    if (group_front(ctx))
        for (int r = 1; r < ctx->group->world; r++) {
            const int32_t rc = zb_sync(ctx->group->child[r]);
            if (rc) return rc;
        }
    tail_quiesce(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return ZB_OK;
}

int32_t zb_set_option(zb_ctx *ctx, const char *key, int64_t value) {
    if (group_front(ctx)) // the same setting on every rank's context
        for (int r = 1; r < ctx->group->world; r++) {
            const int32_t rc = zb_set_option(ctx->group->child[r], key, value);
            if (rc) return rc;
        }
    tail_quiesce(ctx);
    if (key && !strcmp(key, "tail_log2")) {
        if (value < 0 || value > 20) return ZB_ERR_BAD_ARGUMENT;
        ctx->tail_log2 = (int)value;
        return ZB_OK;
    }
    if (key && !strcmp(key, "prelaunch")) {
        ctx->prelaunch = value != 0;
        return ZB_OK;
    }
    if (key && !strcmp(key, "tail_test_starve")) {
        ctx->tail_test_starve = value != 0;
        return ZB_OK;
    }
    if (key && !strcmp(key, "lasso_chunk_log2")) {
        if (value < 4 || value > 30) return ZB_ERR_BAD_ARGUMENT;
        ctx->lasso_chunk_log2 = (int)value;
        return ZB_OK;
    }
    if (key && !strcmp(key, "starved")) {
        ctx->starved = (int)value;
        return ZB_OK;
    }
    if (key && !strcmp(key, "xchg_stats_reset")) {
        if (ctx->d_xchg_stats) CK(cudaMemset(ctx->d_xchg_stats, 0, 2 * sizeof(unsigned long long)));
        return ZB_OK;
    }
    if (key && !strcmp(key, "linear_d1")) {
        ctx->linear_d1 = value != 0;
        return ZB_OK;
    }
    if (key && !strcmp(key, "host_tail_log2")) {
        if (value < 2 || value > LIN_DUMP_MAX_LOG2) return ZB_ERR_BAD_ARGUMENT;
        ctx->host_tail_log2 = (int)value;
        return ZB_OK;
    }
    if (key && !strcmp(key, "prod_host_tail_log2")) {
        if (value < 0 || value > PROD_DUMP_MAX_LOG2) return ZB_ERR_BAD_ARGUMENT;
        ctx->prod_host_tail_log2 = (int)value;
        return ZB_OK;
    }
    if (key && !strcmp(key, "linear_k")) {
        if (value < 1 || value > LIN_MAX_K) return ZB_ERR_BAD_ARGUMENT;
        ctx->linear_k = (int)value;
        return ZB_OK;
    }
    if (key && !strcmp(key, "comm_reduce")) {
        if (value < 0 || value > 2 || (value == 2 && !ctx->d_xchg_view)) return ZB_ERR_BAD_ARGUMENT;
        ctx->comm_reduce = (int)value;
        return ZB_OK;
    }
    return ZB_ERR_BAD_ARGUMENT;
}

int32_t zb_get_option(zb_ctx *ctx, const char *key, int64_t *value) {
    if (key && value && !strcmp(key, "tail_log2")) {
        *value = ctx->tail_log2;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "comm_reduce")) {
        *value = ctx->comm_reduce;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "prelaunch")) {
        *value = ctx->prelaunch ? 1 : 0;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "starved")) {
        *value = ctx->starved;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "linear_d1")) {
        *value = ctx->linear_d1 ? 1 : 0;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "prod_host_tail_log2")) {
        *value = ctx->prod_host_tail_log2;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "host_tail_log2")) {
        *value = ctx->host_tail_log2;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "linear_k")) {
        *value = ctx->linear_k;
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "h2d_bytes")) {
        *value = (int64_t)ctx->h2d_bytes;
        return ZB_OK;
    }
    if (key && value && (!strcmp(key, "xchg_wait_ns") || !strcmp(key, "xchg_rounds"))) {
        // in-kernel peer exchange: time the last CTA spent waiting for the other ranks' rows (skew between GPUs + NVLink
        // latency), summed over the exchanged rounds since the last reset, and the number of those rounds
        unsigned long long st[2] = {0, 0};
        if (ctx->d_xchg_stats) {
            tail_quiesce(ctx);
            CK(cudaMemcpy(st, ctx->d_xchg_stats, sizeof(st), cudaMemcpyDeviceToHost));
        }
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        *value = !strcmp(key, "xchg_rounds") ? (int64_t)st[1] : (int64_t)((double)st[0] * 1e6 / (khz > 0 ? khz : 1965000));
        return ZB_OK;
    }
    if (key && value && !strcmp(key, "p2p_attached")) {
        *value = ctx->d_xchg_view ? 1 : 0;
        return ZB_OK;
    }
    return ZB_ERR_BAD_ARGUMENT;
}

int32_t zb_timer_start(zb_ctx *ctx) {
    tail_quiesce(ctx);
    if (!ctx->t0) {
        CK(cudaEventCreate(&ctx->t0));
        CK(cudaEventCreate(&ctx->t1));
    }
    CK(cudaEventRecord(ctx->t0, ctx->stream));
    return ZB_OK;
}

int32_t zb_timer_stop(zb_ctx *ctx, float *ms) {
    tail_quiesce(ctx);
    if (!ctx->t0 || !ms) return ZB_ERR_BAD_ARGUMENT;
    CK(cudaEventRecord(ctx->t1, ctx->stream));
    CK(cudaEventSynchronize(ctx->t1));
    CK(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return ZB_OK;
}

int32_t zb_profile_enable(zb_ctx *ctx, int32_t on) {
    tail_quiesce(ctx);
    prof_drain(ctx, true);
    if (on) ctx->prof.clear();
    ctx->profiling = on != 0;
    return ZB_OK;
}

uint32_t zb_profile_count(zb_ctx *ctx) {
    tail_quiesce(ctx);
    prof_drain(ctx, true);
    return (uint32_t)ctx->prof.size();
}

int32_t zb_profile_entry(zb_ctx *ctx, uint32_t i, char *name, uint32_t cap, uint64_t *launches, double *ms, uint64_t *bytes) {
    if (i >= ctx->prof.size()) return ZB_ERR_BAD_ARGUMENT;
    const auto &e = ctx->prof[i];
    if (name && cap) snprintf(name, cap, "%s", e.name.c_str());
    if (launches) *launches = e.launches;
    if (ms) *ms = e.ms;
    if (bytes) *bytes = e.bytes;
    return ZB_OK;
}

int32_t zb_int_pipe_peak(zb_ctx *ctx, double *lop3, double *shf, double *mix) {
    tail_quiesce(ctx);
    BufRef buf;
    int32_t rc = dev_alloc(ctx, 256, &buf);
    if (rc) return rc;
    const int ctas = ctx->sm_count * 8, iters = 4096;
    double *outs[3] = {lop3, shf, mix};
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) { // first repetition warms up
            CK(cudaEventRecord(a, ctx->stream));
            launch_int_peak(mode, (uint32_t *)buf->ptr, iters, ctas, ctx->stream);
            ctx->launches++;
            CK(cudaEventRecord(b, ctx->stream));
            CK(cudaEventSynchronize(b));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (rep && ms < best) best = ms;
        }
        const double ops = (double)ctas * 256.0 * iters * 64.0;
        if (outs[mode]) *outs[mode] = ops / (best * 1e-3);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ZB_OK;
}

int32_t zb_h2d_rate(zb_ctx *ctx, size_t bytes, double *bytes_per_s) {
    tail_quiesce(ctx);
    if (!bytes_per_s || bytes < (1u << 20)) return ZB_ERR_BAD_ARGUMENT;
    void *h = nullptr;
    BufRef d;
    int32_t rc = dev_alloc(ctx, bytes, &d);
    if (rc) return rc;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
    memset(h, 1, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) { // the first repetition warms up
        cudaEventRecord(a, ctx->stream);
        cudaMemcpyAsync(d->ptr, h, bytes, cudaMemcpyHostToDevice, ctx->stream);
        cudaEventRecord(b, ctx->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFreeHost(h);
    CK(cudaGetLastError());
    *bytes_per_s = (double)bytes / (best * 1e-3);
    return ZB_OK;
}

int32_t zb_host_alloc(zb_ctx *ctx, size_t bytes, void **out) {
    tail_quiesce(ctx);
    CK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return ZB_OK;
}
int32_t zb_host_free(zb_ctx *ctx, void *p) {
    tail_quiesce(ctx);
    CK(cudaFreeHost(p));
    return ZB_OK;
}

int32_t zb_device_info(zb_ctx *ctx, int32_t *sm_count, uint64_t *total_mem, uint64_t *free_mem) {
    tail_quiesce(ctx);
    size_t f = 0, t = 0;
    CK(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx->sm_count;
    if (total_mem) *total_mem = t;
    if (free_mem) *free_mem = f + ctx->cached_bytes;
    return ZB_OK;
}

/* ------------------------------------------------------------------ Multilinear */

int32_t zb_mle_upload(zb_ctx *ctx, const uint64_t *evals, uint64_t n, zb_mle *out) {
    if (group_front(ctx)) return zg_upload<uint64_t>(ctx, evals, n, out);
    tail_quiesce(ctx);
    int32_t rc = check_pow2(n);
    if (rc) return rc;
    if (!evals || !out) return ZB_ERR_BAD_ARGUMENT;
    Mle *m = nullptr;
    rc = new_mle(ctx, n, out, &m);
    if (rc) return rc;
    rc = upload_narrow(ctx, evals, n, m->d());
    if (rc) {
        ctx->mles.erase(*out);
        *out = 0;
    }
    return rc;
}

int32_t zb_mle_upload_u32(zb_ctx *ctx, const uint32_t *evals, uint64_t n, zb_mle *out) {
    if (group_front(ctx)) return zg_upload<uint32_t>(ctx, evals, n, out);
    tail_quiesce(ctx);
    int32_t rc = check_pow2(n);
    if (rc) return rc;
    if (!evals || !out) return ZB_ERR_BAD_ARGUMENT;
    Mle *m = nullptr;
    rc = new_mle(ctx, n, out, &m);
    if (rc) return rc;
    rc = staged_h2d(ctx, m->d(), evals, n * sizeof(uint32_t));
    if (rc) {
        ctx->mles.erase(*out);
        *out = 0;
        return rc;
    }
    {
        ProfScope _ps(ctx, "check_u32", n * 4);
        launch_check_u32(m->d(), n, ctx->d_err, ctx->stream);
    }
    LAUNCHED("check");
    rc = read_err_flag(ctx);
    if (rc) {
        ctx->mles.erase(*out);
        *out = 0;
    }
    return rc;
}

int32_t zb_mle_constant(zb_ctx *ctx, uint32_t num_vars, uint64_t value, zb_mle *out) {
    if (group_front(ctx)) {
        uint32_t k = 0;
        while ((1 << k) < ctx->group->world) k++;
        if (num_vars < 2 * k || num_vars > 40 || !out) return ZB_ERR_BAD_ARGUMENT;
        return group_make_mle(ctx, 1ull << num_vars, out,
                              [&](int, zb_ctx *c, zb_mle *h) { return zb_mle_constant(c, num_vars - k, value, h); });
    }
    tail_quiesce(ctx);
    if (num_vars > 40 || value >= bb::P || !out) return value >= bb::P ? ZB_ERR_NOT_CANONICAL : ZB_ERR_BAD_ARGUMENT;
    Mle *m = nullptr;
    int32_t rc = new_mle(ctx, 1ull << num_vars, out, &m);
    if (rc) return rc;
    {
        ProfScope _ps(ctx, "fill", m->n * 4);
        launch_fill(m->d(), m->n, (uint32_t)value, ctx->stream);
    }
    LAUNCHED("fill");
    return zb_sync(ctx);
}

int32_t zb_mle_synthetic(zb_ctx *ctx, uint64_t seed, uint64_t start, uint64_t stride, uint64_t n, zb_mle *out) {
    if (group_front(ctx)) { // rank r generates its cyclic shard: global elements r, r + P, r + 2 P, ...
        const uint64_t P = (uint64_t)ctx->group->world;
        if (check_pow2(n) || n < P * P || !out) return check_pow2(n) ? check_pow2(n) : ZB_ERR_BAD_ARGUMENT;
        return group_make_mle(ctx, n, out, [&](int r, zb_ctx *c, zb_mle *h) {
            return zb_mle_synthetic(c, seed, start + (uint64_t)r * stride, stride * P, n / P, h);
        });
    }
    tail_quiesce(ctx);
    int32_t rc = check_pow2(n);
    if (rc) return rc;
    Mle *m = nullptr;
    rc = new_mle(ctx, n, out, &m);
    if (rc) return rc;
    {
        ProfScope _ps(ctx, "synthetic", n * 4);
        launch_synthetic(m->d(), n, seed, start, stride, ctx->stream);
    }
    LAUNCHED("synthetic");
    return zb_sync(ctx);
}

int32_t zb_mle_clone(zb_ctx *ctx, zb_mle src, zb_mle *out) {
    tail_quiesce(ctx);
    Mle *s = get_mle(ctx, src);
    if (!s) return ZB_ERR_BAD_HANDLE;
    uint64_t n = s->n;
    BufRef sb = s->buf;
    Mle *m = nullptr;
    int32_t rc = new_mle(ctx, n, out, &m);
    if (rc) return rc;
    const cudaError_t ce = cudaMemcpyAsync(m->d(), sb->ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
    rc = ce == cudaSuccess ? zb_sync(ctx) : cuda_fail(ctx, ce, "cudaMemcpyAsync(clone)");
    if (rc) {
        ctx->mles.erase(*out);
        *out = 0;
    }
    return rc;
}

int32_t zb_mle_free(zb_ctx *ctx, zb_mle h) {
    if (group_front(ctx) && is_group_handle(h)) return zg_mle_free(ctx, h);
    tail_quiesce(ctx);
    if (!ctx->mles.erase(h)) return ZB_ERR_BAD_HANDLE;
    return ZB_OK;
}

int32_t zb_mle_len(zb_ctx *ctx, zb_mle h, uint64_t *n, uint32_t *num_vars) {
    if (group_front(ctx) && is_group_handle(h)) {
        GMle *gm = get_gmle(ctx, h);
        if (!gm) return ZB_ERR_BAD_HANDLE;
        if (n) *n = gm->n;
        if (num_vars) *num_vars = (uint32_t)__builtin_ctzll(gm->n);
        return ZB_OK;
    }
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (n) *n = m->n;
    if (num_vars) *num_vars = (uint32_t)__builtin_ctzll(m->n);
    return ZB_OK;
}

int32_t zb_mle_download(zb_ctx *ctx, zb_mle h, uint64_t *out, uint64_t n) { return zb_mle_download_range(ctx, h, 0, out, n); }

int32_t zb_mle_download_range(zb_ctx *ctx, zb_mle h, uint64_t offset, uint64_t *out, uint64_t n) {
    if (group_front(ctx) && is_group_handle(h)) return zg_mle_download(ctx, h, offset, out, n);
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (offset > m->n || n > m->n - offset || !out) return ZB_ERR_BAD_ARGUMENT;
    uint64_t chunk = n < STAGE_ELEMS ? n : STAGE_ELEMS;
    BufRef stage;
    int32_t rc = dev_alloc(ctx, chunk * sizeof(uint64_t), &stage);
    if (rc) return rc;
    for (uint64_t off = 0; off < n; off += chunk) {
        uint64_t k = n - off < chunk ? n - off : chunk;
        {
            ProfScope _ps(ctx, "widen_u32", k * 12);
            launch_widen_u32(m->d() + offset + off, (uint64_t *)stage->ptr, k, ctx->stream);
        }
        LAUNCHED("widen");
        CK(cudaMemcpyAsync(out + off, stage->ptr, k * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return zb_sync(ctx);
}

int32_t zb_mle_download_u32(zb_ctx *ctx, zb_mle h, uint64_t offset, uint32_t *out, uint64_t n) {
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (offset > m->n || n > m->n - offset || !out) return ZB_ERR_BAD_ARGUMENT;
    CK(cudaMemcpyAsync(out, m->d() + offset, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return zb_sync(ctx);
}

int32_t zb_host_scratch(zb_ctx *ctx, size_t bytes, void **out) {
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    if (bytes > ctx->scratch_bytes) {
        tail_quiesce(ctx);
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) cudaFreeHost(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        CK(cudaHostAlloc(&ctx->scratch, bytes, cudaHostAllocDefault));
        ctx->scratch_bytes = bytes;
    }
    *out = ctx->scratch;
    return ZB_OK;
}

int32_t zb_host_mirror(zb_ctx *ctx, size_t bytes, void **out) {
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    if (bytes > ctx->mirror_bytes) {
        tail_quiesce(ctx);
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->mirror) cudaFreeHost(ctx->mirror);
        ctx->mirror = nullptr;
        ctx->mirror_bytes = 0;
        CK(cudaHostAlloc(&ctx->mirror, bytes, cudaHostAllocDefault));
        ctx->mirror_bytes = bytes;
    }
    *out = ctx->mirror;
    return ZB_OK;
}

int32_t zb_mle_sum(zb_ctx *ctx, zb_mle h, uint64_t *out) {
    if (group_front(ctx) && is_group_handle(h)) return zg_mle_sum(ctx, h, out);
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    Mailbox mb = ctx->mailbox();
    {
        ProfScope _ps(ctx, "sum", m->n * 4);
        launch_sum(m->d(), m->n, mb, ctx->sm_count, ctx->stream);
    }
    LAUNCHED("sum");
    int32_t rc = wait_mail(ctx, mb.seq, 1);
    if (rc) return rc;
    *out = mail_word(ctx, 0);
    return ZB_OK;
}

int32_t zb_mle_round_sums(zb_ctx *ctx, zb_mle h, uint64_t out[2]) {
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (m->n < 2) return ZB_ERR_NO_VARIABLES; // multilinear.zig:207
    PolySet ps{};
    ps.src[0] = m->d();
    Mailbox mb = ctx->mailbox();
    {
        ProfScope _ps(ctx, "round_sums_d1", m->n * 4);
        launch_round_sums(1, ps, m->n, mb, ctx->sm_count, ctx->stream);
    }
    LAUNCHED("round_sums");
    int32_t rc = wait_mail(ctx, mb.seq, 2);
    if (rc) return rc;
    out[0] = mail_word(ctx, 0);
    out[1] = mail_word(ctx, 1);
    return ZB_OK;
}

static int32_t fold_common(zb_ctx *ctx, const uint32_t *src, uint32_t *dst, uint64_t n, uint64_t r, uint64_t next[2]) {
    if (r >= bb::P) return ZB_ERR_NOT_CANONICAL;
    PolySet ps{};
    ps.src[0] = src;
    ps.dst[0] = dst;
    Mailbox mb = ctx->mailbox();
    {
        ProfScope _ps(ctx, "fold_sums_d1", n * 6);
        launch_fold_sums(1, ps, n, (uint32_t)r, mb, ctx->sm_count, ctx->stream);
    }
    LAUNCHED("fold_sums");
    int32_t rc = wait_mail(ctx, mb.seq, n == 2 ? 1 : 2);
    if (rc) return rc;
    if (next) {
        next[0] = mail_word(ctx, 0);
        next[1] = n == 2 ? 0 : mail_word(ctx, 1);
    }
    return ZB_OK;
}

int32_t zb_mle_partial_eval(zb_ctx *ctx, zb_mle h, uint64_t r, zb_mle *out, uint64_t next[2]) {
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (m->n < 2) return ZB_ERR_NO_VARIABLES; // multilinear.zig:156
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    uint64_t n = m->n;
    BufRef src = m->buf; // keep alive across the map insertion below
    Mle *d = nullptr;
    int32_t rc = new_mle(ctx, n / 2, out, &d);
    if (rc) return rc;
    rc = fold_common(ctx, (const uint32_t *)src->ptr, d->d(), n, r, next);
    if (rc) {
        ctx->mles.erase(*out);
        *out = 0;
    }
    return rc;
}

static int32_t fold_inplace_impl(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t r, unsigned long long payload[4],
                                 bool *was_last);

int32_t zb_mle_fold_inplace(zb_ctx *ctx, zb_mle h, uint64_t r, uint64_t next[2]) {
    unsigned long long payload[4];
    bool last = false;
    int32_t rc = fold_inplace_impl(ctx, &h, 1, r, payload, &last);
    if (rc) return rc;
    if (next) {
        next[0] = payload[0];
        next[1] = last ? 0 : payload[1];
    }
    return ZB_OK;
}

int32_t zb_mle_eval(zb_ctx *ctx, zb_mle h, const uint64_t *point, uint32_t npoint, uint64_t *out) {
    if (group_front(ctx) && is_group_handle(h)) return zg_mle_eval(ctx, h, point, npoint, out);
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    uint32_t v = (uint32_t)__builtin_ctzll(m->n);
    if (npoint != v) return ZB_ERR_WRONG_NUM_VARS; // multilinear.zig:112
    for (uint32_t i = 0; i < v; i++)
        if (point[i] >= bb::P) return ZB_ERR_NOT_CANONICAL;
    const uint32_t *src = m->d();
    uint64_t n = m->n;
    BufRef scratch[2];
    int which = 0;
    uint32_t done = 0;
    Mailbox mb = ctx->mailbox();
    // big stages fold 10 variables per pass with the warp-autonomous kernel; the remainder (< 2^20 elements) goes
    // through the block kernel, up to 12 variables per pass; the last pass publishes the value
    do {
        const bool big = (n >= (1ull << 20));
        if (!big && v - done <= 20) { // everything that is left in ONE launch (tiles per CTA, then the last CTA finishes)
            const int nv = (int)(v - done < 12 ? v - done : 12), nv2 = (int)(v - done) - nv;
            EvalPoint pt{}, pt2{};
            for (int k = 0; k < nv; k++) pt.r[k] = (uint32_t)point[done + k], pt.rp[k] = bb::shoup_pre(pt.r[k]);
            for (int k = 0; k < nv2; k++) pt2.r[k] = (uint32_t)point[done + nv + k], pt2.rp[k] = bb::shoup_pre(pt2.r[k]);
            const uint64_t n_out = n >> nv;
            int32_t rc = dev_alloc(ctx, (n_out < 128 ? 128 : n_out) * sizeof(uint32_t), &scratch[which]);
            if (rc) return rc;
            {
                ProfScope _ps(ctx, "eval_finish", (n + n_out) * 4);
                launch_eval_finish(src, n, nv, pt, (uint32_t *)scratch[which]->ptr, nv2, pt2, mb, ctx->sm_count, ctx->stream);
            }
            LAUNCHED("eval_finish");
            break;
        }
        int nv = big ? 10 : (int)(v - done < 12 ? v - done : 12);
        EvalPoint pt{};
        for (int k = 0; k < nv; k++) {
            pt.r[k] = (uint32_t)point[done + k];
            pt.rp[k] = bb::shoup_pre(pt.r[k]);
        }
        uint64_t n_out = n >> nv;
        int32_t rc = dev_alloc(ctx, (n_out < 128 ? 128 : n_out) * sizeof(uint32_t), &scratch[which]);
        if (rc) return rc;
        bool last = (done + nv == v);
        {
            ProfScope _ps(ctx, big ? "eval_warp10" : "eval_stage", (n + n_out) * 4);
            if (big)
                launch_eval_warp10(src, n, pt, (uint32_t *)scratch[which]->ptr, ctx->sm_count, ctx->stream);
            else
                launch_eval_stage(src, n, nv, pt, (uint32_t *)scratch[which]->ptr, last ? &mb : nullptr, ctx->sm_count, ctx->stream);
        }
        LAUNCHED("eval_stage");
        src = (const uint32_t *)scratch[which]->ptr;
        n = n_out;
        done += nv;
        which ^= 1;
    } while (done < v);
    int32_t rc = wait_mail(ctx, mb.seq, 1);
    if (rc) return rc;
    *out = mail_word(ctx, 0);
    return ZB_OK;
}

// Multilinear.eval for `count` polynomials of equal length, one point each, in one launch set and ONE read-back
// (Prover.generateCommitments derives all 43 opening points before it absorbs any value, prover.zig:420-443).
int32_t zb_mle_eval_batch(zb_ctx *ctx, const zb_mle *polys, uint32_t count, const uint64_t *points, uint32_t npoint, uint64_t *out) {
    tail_quiesce(ctx);
    if (!polys || !out || count == 0 || count > (DUMP_BYTES / sizeof(unsigned long long)) || (npoint && !points)) return ZB_ERR_BAD_ARGUMENT;
    std::vector<const uint32_t *> ptrs(count);
    uint64_t n = 0;
    for (uint32_t i = 0; i < count; i++) {
        Mle *m = get_mle(ctx, polys[i]);
        if (!m) return ZB_ERR_BAD_HANDLE;
        if (i == 0) n = m->n;
        else if (m->n != n) return ZB_ERR_DIFFERENT_NUM_VARS;
        ptrs[i] = m->d();
    }
    const uint32_t v = (uint32_t)__builtin_ctzll(n);
    if (npoint != v) return ZB_ERR_WRONG_NUM_VARS; // multilinear.zig:112
    for (uint64_t i = 0; i < (uint64_t)count * v; i++)
        if (points[i] >= bb::P) return ZB_ERR_NOT_CANONICAL;
    // one parameter blob: [count table pointers][count x (v challenges, v Shoup companions)]
    const uint32_t vv = v ? v : 1;
    const size_t blob_bytes = (size_t)count * 8 + (size_t)count * 2 * vv * 4;
    uint8_t *h_blob = nullptr;
    int32_t rc = zb_host_scratch(ctx, blob_bytes, (void **)&h_blob);
    if (rc) return rc;
    memcpy(h_blob, ptrs.data(), (size_t)count * 8);
    uint32_t *hp = (uint32_t *)(h_blob + (size_t)count * 8);
    for (uint32_t i = 0; i < count; i++)
        for (uint32_t k = 0; k < v; k++) {
            const uint32_t r = (uint32_t)points[(size_t)i * v + k];
            hp[(size_t)i * 2 * vv + k] = r;
            hp[(size_t)i * 2 * vv + vv + k] = bb::shoup_pre(r);
        }
    BufRef d_blob, scratch[2];
    rc = dev_alloc(ctx, blob_bytes, &d_blob);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d_blob->ptr, h_blob, blob_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t *const *d_ptrs = (const uint32_t *const *)d_blob->ptr;
    const uint32_t *d_pts = (const uint32_t *)((uint8_t *)d_blob->ptr + (size_t)count * 8);
    unsigned long long *d_dump = (unsigned long long *)((uint8_t *)ctx->d_mail + DUMP_OFFSET);
    Mailbox mb = ctx->mailbox();
    uint64_t cur_n = n;
    uint32_t done = 0;
    int which = 0;
    const uint32_t *rows = nullptr; // after the first stage the partial results of all polynomials are rows of one buffer
    for (;;) {
        const bool big = cur_n >= (1ull << 20);
        const int nv = big ? 10 : (int)(v - done < 12 ? v - done : 12);
        const uint64_t n_out = cur_n >> nv;
        const bool last = done + nv == v;
        rc = dev_alloc(ctx, (size_t)count * (n_out < 32 ? 32 : n_out) * sizeof(uint32_t), &scratch[which]);
        if (rc) return rc;
        {
            ProfScope _ps(ctx, big ? "eval_warp10_batch" : "eval_stage_batch", (uint64_t)count * (cur_n + n_out) * 4);
            if (big && rows == nullptr)
                launch_eval_warp10_batch(d_ptrs, cur_n, count, d_pts, vv, done, (uint32_t *)scratch[which]->ptr, ctx->sm_count, ctx->stream);
            else
                launch_eval_stage_batch(rows ? nullptr : d_ptrs, rows, cur_n, count, nv, d_pts, vv, done, (uint32_t *)scratch[which]->ptr,
                                        last ? d_dump : nullptr, mb, ctx->stream);
        }
        LAUNCHED("eval_batch");
        rows = (const uint32_t *)scratch[which]->ptr;
        cur_n = n_out;
        done += nv;
        which ^= 1;
        if (last) break;
    }
    rc = wait_mail(ctx, mb.seq);
    if (rc) return rc;
    const unsigned long long *res = (const unsigned long long *)((uint8_t *)ctx->h_mail + DUMP_OFFSET);
    for (uint32_t i = 0; i < count; i++) out[i] = res[i];
    return ZB_OK;
}

// eq(tau, .) as a table (extension: weights of an eq-weighted sumcheck). E[i] = prod_k (bit_k(i) ? tau[k] : 1 - tau[k]) — index bit
// k <-> tau[k], the convention of Multilinear.eval (multilinear.zig:128-141), so sum_i E[i] A[i] == A.eval(tau).
int32_t zb_mle_eq(zb_ctx *ctx, const uint64_t *tau, uint32_t v, zb_mle *out) {
    tail_quiesce(ctx);
    if (!out || v > 40 || (v && !tau)) return ZB_ERR_BAD_ARGUMENT;
    for (uint32_t k = 0; k < v; k++)
        if (tau[k] >= bb::P) return ZB_ERR_NOT_CANONICAL;
    const uint32_t s = v / 2, t = v - s; // low s variables / high t variables (t <= 20)
    auto half = [&](uint32_t k0, uint32_t cnt, bool mont) {
        std::vector<uint32_t> e(1ull << cnt);
        e[0] = mont ? bb::R_MOD_P : 1u;
        for (uint32_t j = 0; j < cnt; j++) { // after step j: entries for the bits 0..j of this half
            const uint32_t r = (uint32_t)tau[k0 + j], nr = bb::sub(1u, r);
            for (uint64_t i = 0; i < (1ull << j); i++) {
                const uint32_t base = e[i];
                e[i] = bb::mul(base, nr);
                e[i + (1ull << j)] = bb::mul(base, r);
            }
        }
        return e;
    };
    const std::vector<uint32_t> lo = half(0, s, false), hi = half(s, t, true);
    BufRef dlo, dhi;
    int32_t rc = dev_alloc(ctx, lo.size() * 4, &dlo);
    if (rc == ZB_OK) rc = dev_alloc(ctx, hi.size() * 4, &dhi);
    if (rc) return rc;
    Mle *m = nullptr;
    rc = new_mle(ctx, 1ull << v, out, &m);
    if (rc) return rc;
    cudaError_t ce = cudaMemcpyAsync(dlo->ptr, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(dhi->ptr, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) {
        ProfScope _ps(ctx, "eq_table", 4ull << v);
        launch_eq_table((const uint32_t *)dlo->ptr, (const uint32_t *)dhi->ptr, (int)s, 1ull << v, m->d(), ctx->stream);
    }
    rc = ce == cudaSuccess ? check_launch(ctx, "eq_table") : cuda_fail(ctx, ce, "cudaMemcpyAsync(eq halves)");
    if (rc == ZB_OK) rc = zb_sync(ctx); // lo / hi are stack-owned host vectors
    else cudaStreamSynchronize(ctx->stream);
    if (rc) {
        ctx->mles.erase(*out);
        *out = 0;
    }
    return rc;
}

int32_t zb_mle_add(zb_ctx *ctx, zb_mle a, zb_mle b, zb_mle *out) {
    tail_quiesce(ctx);
    Mle *ma = get_mle(ctx, a), *mb_ = get_mle(ctx, b);
    if (!ma || !mb_) return ZB_ERR_BAD_HANDLE;
    if (ma->n != mb_->n) return ZB_ERR_DIFFERENT_NUM_VARS; // multilinear.zig:237
    BufRef ba = ma->buf, bbuf = mb_->buf;
    uint64_t n = ma->n;
    Mle *o = nullptr;
    int32_t rc = new_mle(ctx, n, out, &o);
    if (rc) return rc;
    {
        ProfScope _ps(ctx, "add", n * 12);
        launch_add((const uint32_t *)ba->ptr, (const uint32_t *)bbuf->ptr, o->d(), n, ctx->stream);
    }
    LAUNCHED("add");
    return zb_sync(ctx);
}

int32_t zb_mle_scalar_mul(zb_ctx *ctx, zb_mle a, uint64_t scalar, zb_mle *out) {
    tail_quiesce(ctx);
    Mle *ma = get_mle(ctx, a);
    if (!ma) return ZB_ERR_BAD_HANDLE;
    if (scalar >= bb::P) return ZB_ERR_NOT_CANONICAL;
    BufRef ba = ma->buf;
    uint64_t n = ma->n;
    Mle *o = nullptr;
    int32_t rc = new_mle(ctx, n, out, &o);
    if (rc) return rc;
    {
        ProfScope _ps(ctx, "scalar_mul", n * 8);
        launch_scalar_mul((const uint32_t *)ba->ptr, (uint32_t)scalar, o->d(), n, ctx->stream);
    }
    LAUNCHED("scalar_mul");
    return zb_sync(ctx);
}

/* ------------------------------------------------------------------ product sumcheck rounds */

static inline uint32_t f_add(uint32_t a, uint32_t b) { return bb::add(a, b); }
static inline uint32_t f_sub(uint32_t a, uint32_t b) { return bb::sub(a, b); }
static inline uint32_t f_half(uint32_t a) { return bb::mul(a, (bb::P + 1) / 2); }

// mailbox payload (evaluations) -> coefficient form [a0..ad]
static void evals_to_coeffs(uint32_t d, const unsigned long long *mail, uint64_t *out) {
    if (d == 1) {
        uint32_t s0 = (uint32_t)mail[0], s1 = (uint32_t)mail[1];
        out[0] = s0;
        out[1] = f_sub(s1, s0); // multilinear.zig:229
    } else if (d == 2) {
        uint32_t g0 = (uint32_t)mail[0], g1 = (uint32_t)mail[1], gi = (uint32_t)mail[2];
        out[0] = g0;
        out[2] = gi;
        out[1] = f_sub(f_sub(g1, g0), gi);
    } else {
        uint32_t g0 = (uint32_t)mail[0], g1 = (uint32_t)mail[1], gm = (uint32_t)mail[2], gi = (uint32_t)mail[3];
        out[0] = g0;
        out[3] = gi;
        out[2] = f_sub(f_half(f_add(g1, gm)), g0);
        out[1] = f_sub(f_half(f_sub(g1, gm)), gi);
    }
}

static int32_t gather_polys(zb_ctx *ctx, const zb_mle *polys, uint32_t d, Mle **ms) {
    if (d < 1 || d > (uint32_t)MAX_POLYS || !polys) return ZB_ERR_BAD_ARGUMENT;
    for (uint32_t k = 0; k < d; k++) {
        ms[k] = get_mle(ctx, polys[k]);
        if (!ms[k]) return ZB_ERR_BAD_HANDLE;
        if (ms[k]->n != ms[0]->n) return ZB_ERR_DIFFERENT_NUM_VARS;
        for (uint32_t j = 0; j < k; j++)
            if (polys[j] == polys[k]) return ZB_ERR_BAD_ARGUMENT;
    }
    return ZB_OK;
}

int32_t zb_prod_round_coeffs(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *out) {
    tail_quiesce(ctx);
    Mle *ms[MAX_POLYS];
    int32_t rc = gather_polys(ctx, polys, d, ms);
    if (rc) return rc;
    if (ms[0]->n < 2) return ZB_ERR_NO_VARIABLES;
    PolySet ps{};
    for (uint32_t k = 0; k < d; k++) ps.src[k] = ms[k]->d();
    const bool red = reduce_on_device(ctx);
    Mailbox mb = round_mailbox(ctx, red);
    {
        ProfScope _ps(ctx, d == 1 ? "round_sums_d1" : d == 2 ? "round_sums_d2" : "round_sums_d3", ms[0]->n * 4 * d);
        launch_round_sums((int)d, ps, ms[0]->n, mb, ctx->sm_count, ctx->stream);
    }
    LAUNCHED("prod_round_sums");
    if (red) {
        rc = comm_publish(ctx, mb.seq, d == 1 ? 2 : (int)d + 1);
        if (rc) return rc;
    }
    rc = wait_mail(ctx, mb.seq, mb.tagged ? (d == 1 ? 2 : (int)d + 1) : 0);
    if (rc) return rc;
    evals_to_coeffs(d, ctx->h_mail, out);
    return ZB_OK;
}

// Waits for mailbox sequence `seq` of a kernel that polls the host for its challenge (tail session or pre-launched
// fold). *gone = true when the kernel has left instead (starvation exit: launches serialised by a profiler, or the host
// thread lost the CPU): the tables are untouched for this round and the caller redoes it with a plain launch.
static int32_t wait_polling_kernel(zb_ctx *ctx, unsigned long long seq, int tagged_words, bool *gone) {
    auto t0 = std::chrono::steady_clock::now();
    uint64_t spins = 0;
    *gone = false;
    while (!mail_ready(ctx, seq, tagged_words)) {
        if ((++spins & 0x3FFF) == 0) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) {
                const int32_t e = cuda_fail(ctx, q, "polling kernel");
                rearm(ctx);
                return e;
            }
            if (q == cudaSuccess) { // drained: give the mapped write 2 ms to land, then decide
                auto t1 = std::chrono::steady_clock::now();
                while (!mail_ready(ctx, seq, tagged_words) && std::chrono::steady_clock::now() - t1 < std::chrono::milliseconds(2)) {
                }
                if (!mail_ready(ctx, seq, tagged_words)) *gone = true;
                break;
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
                tail_quiesce(ctx);
                rearm(ctx);
                ctx->last_error = "timeout waiting for a polling kernel";
                return ZB_ERR_TIMEOUT;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (!*gone && ctx->comm_reduce == 2 && ctx->d_xchg_view && ctx->h_mail[MAIL_WORDS - 1] == 1ull) {
        ctx->h_mail[MAIL_WORDS - 1] = 0;
        ctx->last_error = "peer exchange: a rank never arrived";
        rearm(ctx);
        return ZB_ERR_TIMEOUT;
    }
    return ZB_OK;
}

static void feed_challenge(zb_ctx *ctx, unsigned int tag, uint64_t r) {
    if (ctx->tail_test_starve) { // test hook: delay the challenge past the polling kernel's patience
        std::this_thread::sleep_for(std::chrono::milliseconds(400));
        ctx->tail_test_starve = false;
    }
    __atomic_store_n(ctx->h_chal, ((unsigned long long)tag << 32) | (unsigned long long)r, __ATOMIC_RELEASE);
}

// After the kernel of the current round (tables of length n_now -> n_now / 2) has been enqueued: put the NEXT round's
// kernel behind it right away, so that its launch latency overlaps the current kernel and the host's transcript work.
// Small next tables start the persistent tail session early, big ones get a polling fold kernel (ChalSrc).
static int32_t prelaunch_next(zb_ctx *ctx, const zb_mle *polys, uint32_t d, Mle **ms, uint64_t n_next) {
    if (!ctx->prelaunch || ctx->tail.active || ctx->pre.active || n_next < 4) return ZB_OK;
    const bool red = reduce_on_device(ctx) && n_next > 2;
    if (red && !reduce_p2p(ctx)) return ZB_OK; // the NCCL exchange enqueues its own work between rounds
    PolySet ps{};
    for (uint32_t k = 0; k < d; k++) {
        ps.src[k] = ms[k]->d();
        ps.dst[k] = ms[k]->d();
    }
    if (!red && ctx->tail_log2 > 0 && n_next <= (1ull << ctx->tail_log2)) {
        if (ctx->chal_seq > 0xF0000000u) ctx->chal_seq = 0;
        Mailbox mb = ctx->mailbox();
        ctx->seq--; // the session's first round takes the number the next mailbox() call hands out
        mb.seq = ctx->seq + 1;
        {
            ProfScope _ps(ctx, d == 1 ? "tail_rounds_d1" : d == 2 ? "tail_rounds_d2" : "tail_rounds_d3", 0);
            launch_tail_rounds((int)d, ps, n_next, mb, ctx->d_chal, ctx->chal_seq + 1, ctx->d_tail_status, ctx->stream);
        }
        int32_t rc = check_launch(ctx, "tail_rounds");
        if (rc) return rc;
        ctx->tail.active = true;
        ctx->tail.d = d;
        ctx->tail.n = n_next;
        for (uint32_t k = 0; k < d; k++) ctx->tail.h[k] = polys[k];
        return ZB_OK;
    }
    if (!fold_sums_is_vector(n_next)) return ZB_OK;
    if (ctx->chal_seq > 0xF0000000u) ctx->chal_seq = 0;
    ctx->pre.mb = round_mailbox(ctx, red);
    ctx->pre.tag = ++ctx->chal_seq;
    ctx->pre.red = red;
    ChalSrc cs{ctx->d_chal, ctx->d_bcast, ctx->d_claim, ctx->pre.tag};
    {
        ProfScope _ps(ctx, d == 1 ? "fold_sums_d1" : d == 2 ? "fold_sums_d2" : "fold_sums_d3", n_next * 6 * d);
        launch_fold_sums((int)d, ps, n_next, 0, ctx->pre.mb, ctx->sm_count, ctx->stream, &cs);
    }
    int32_t rc = check_launch(ctx, "fold_sums(pre)");
    if (rc) return rc;
    ctx->pre.active = true;
    ctx->pre.d = d;
    ctx->pre.n = n_next;
    for (uint32_t k = 0; k < d; k++) ctx->pre.h[k] = polys[k];
    return ZB_OK;
}

// One in-place fold round of d polynomials with challenge r; payload = the kernel's raw mailbox words
// (round evaluations of the folded tables, or the d final evaluations when the tables had 2 entries).
// Small tables go through the persistent tail session, everything else through one kernel per round — launched here,
// or already waiting on the device if the previous round pre-launched it.
static int32_t fold_inplace_impl(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t r, unsigned long long payload[4],
                                 bool *was_last) {
    if (d < 1 || d > (uint32_t)MAX_POLYS || !polys) return ZB_ERR_BAD_ARGUMENT;
    bool in_tail = ctx->tail.active && ctx->tail.d == d, in_pre = ctx->pre.active && ctx->pre.d == d;
    for (uint32_t k = 0; k < d; k++) {
        in_tail = in_tail && ctx->tail.h[k] == polys[k];
        in_pre = in_pre && ctx->pre.h[k] == polys[k];
    }
    if (!in_tail && !in_pre) tail_quiesce(ctx);
    else ensure_device(ctx);
    Mle *ms[MAX_POLYS];
    int32_t rc = gather_polys(ctx, polys, d, ms);
    if (rc) return rc;
    const uint64_t n = ms[0]->n;
    if (n < 2) return ZB_ERR_NO_VARIABLES;
    if (r >= bb::P) return ZB_ERR_NOT_CANONICAL;
    if ((in_tail && ctx->tail.n != n) || (in_pre && ctx->pre.n != n)) { // cannot happen through the ABI; be safe
        tail_quiesce(ctx);
        in_tail = in_pre = false;
    }
    *was_last = (n == 2);
    PolySet ps{};
    for (uint32_t k = 0; k < d; k++) {
        ps.src[k] = ms[k]->d();
        ps.dst[k] = ms[k]->d();
    }
    const char *name = d == 1 ? "fold_sums_d1" : d == 2 ? "fold_sums_d2" : "fold_sums_d3";
    const bool red = reduce_on_device(ctx) && n > 2; // the final evaluations (n == 2) are per-rank values, never summed
    Mailbox mb{};
    bool polling = false; // this round's kernel takes its challenge from the host-mapped word
    if (in_pre) {
        mb = ctx->pre.mb;
        ctx->pre.active = false;
        feed_challenge(ctx, ctx->pre.tag, r);
        polling = true;
    } else {
        if (!in_tail && !red && ctx->tail_log2 > 0 && n >= 4 && n <= (1ull << ctx->tail_log2)) {
            // start a tail session now: the kernel serves this and every later round of these tables
            if (ctx->chal_seq > 0xF0000000u) ctx->chal_seq = 0;
            Mailbox tmb = ctx->mailbox();
            ctx->seq--;
            tmb.seq = ctx->seq + 1;
            {
                ProfScope _ps(ctx, d == 1 ? "tail_rounds_d1" : d == 2 ? "tail_rounds_d2" : "tail_rounds_d3", 0);
                launch_tail_rounds((int)d, ps, n, tmb, ctx->d_chal, ctx->chal_seq + 1, ctx->d_tail_status, ctx->stream);
            }
            rc = check_launch(ctx, "tail_rounds");
            if (rc) return rc;
            ctx->tail.active = true;
            ctx->tail.d = d;
            ctx->tail.n = n;
            for (uint32_t k = 0; k < d; k++) ctx->tail.h[k] = polys[k];
            in_tail = true;
        }
        if (in_tail) {
            mb = ctx->mailbox();
            feed_challenge(ctx, ++ctx->chal_seq, r);
            polling = true;
        } else {
            mb = round_mailbox(ctx, red);
            {
                ProfScope _ps(ctx, name, n * 6 * d);
                launch_fold_sums((int)d, ps, n, (uint32_t)r, mb, ctx->sm_count, ctx->stream);
            }
            rc = check_launch(ctx, "fold_sums");
            if (rc) return rc;
            if (red) {
                rc = comm_publish(ctx, mb.seq, d == 1 ? 2 : (int)d + 1);
                if (rc) return rc;
            }
        }
    }
    // the next round's kernel goes behind this one before we wait for this one's sums
    if (!in_tail) {
        rc = prelaunch_next(ctx, polys, d, ms, n / 2);
        if (rc) return rc;
    }
    const int nw = n == 2 ? (int)d : (d == 1 ? 2 : (int)d + 1); // payload words of this round's kernel
    if (polling) {
        bool gone = false;
        rc = wait_polling_kernel(ctx, mb.seq, mb.tagged ? nw : 0, &gone);
        if (rc) return rc;
        if (gone) {
            // redo the round with a plain launch (same exchange round number) and stop using polling kernels here
            tail_quiesce(ctx); // also ends a kernel that was pre-launched behind the one that left
            if (++ctx->starved >= 3) { // once may be a hiccup of the host thread; repeatedly means launches are serialised
                ctx->tail_log2 = 0;
                ctx->prelaunch = false;
            }
            cudaMemsetAsync(ctx->d_tail_status, 0, sizeof(unsigned int), ctx->stream);
            Mailbox mb2 = ctx->mailbox();
            mb2.xchg = mb.xchg;
            mb2.xseq = mb.xseq;
            mb2.tagged = !red;
            if (red && !mb2.xchg) mb2.mail = ctx->d_comm;
            {
                ProfScope _ps(ctx, name, n * 6 * d);
                launch_fold_sums((int)d, ps, n, (uint32_t)r, mb2, ctx->sm_count, ctx->stream);
            }
            rc = check_launch(ctx, "fold_sums");
            if (rc) return rc;
            if (red) {
                rc = comm_publish(ctx, mb2.seq, d == 1 ? 2 : (int)d + 1);
                if (rc) return rc;
            }
            rc = wait_mail(ctx, mb2.seq, mb2.tagged ? nw : 0);
            if (rc) return rc;
        } else if (in_tail) {
            ctx->tail.n = n / 2;
            if (n == 2) ctx->tail.active = false; // the kernel returns after the last round
        }
    } else {
        rc = wait_mail(ctx, mb.seq, mb.tagged ? nw : 0);
        if (rc) return rc;
    }
    for (uint32_t k = 0; k < d; k++) ms[k]->n = n / 2;
    for (int k = 0; k < 4; k++) payload[k] = mail_word(ctx, k);
    return ZB_OK;
}

int32_t zb_prod_fold_inplace(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t r, uint64_t *next) {
    unsigned long long payload[4];
    bool last = false;
    int32_t rc = fold_inplace_impl(ctx, polys, d, r, payload, &last);
    if (rc) return rc;
    if (next) {
        if (last)
            for (uint32_t k = 0; k < d; k++) next[k] = payload[k];
        else
            evals_to_coeffs(d, payload, next);
    }
    return ZB_OK;
}

int32_t zb_prod_partial_eval(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t r, zb_mle *out, uint64_t *next) {
    tail_quiesce(ctx);
    Mle *ms[MAX_POLYS];
    int32_t rc = gather_polys(ctx, polys, d, ms);
    if (rc) return rc;
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    uint64_t n = ms[0]->n;
    if (n < 2) return ZB_ERR_NO_VARIABLES; // multilinear.zig:156
    if (r >= bb::P) return ZB_ERR_NOT_CANONICAL;
    PolySet ps{};
    BufRef keep[MAX_POLYS];
    for (uint32_t k = 0; k < d; k++) {
        ps.src[k] = ms[k]->d();
        keep[k] = ms[k]->buf;
    }
    for (uint32_t k = 0; k < d; k++) { // allocation may rehash the handle table: ms[] is dead from here on
        Mle *o = nullptr;
        rc = new_mle(ctx, n / 2, &out[k], &o);
        if (rc) {
            for (uint32_t j = 0; j < k; j++) ctx->mles.erase(out[j]);
            return rc;
        }
        ps.dst[k] = o->d();
    }
    const bool red = reduce_on_device(ctx) && n > 2;
    Mailbox mb = round_mailbox(ctx, red);
    {
        ProfScope _ps(ctx, d == 1 ? "fold_sums_d1" : d == 2 ? "fold_sums_d2" : "fold_sums_d3", n * 6 * d);
        launch_fold_sums((int)d, ps, n, (uint32_t)r, mb, ctx->sm_count, ctx->stream);
    }
    rc = check_launch(ctx, "prod_partial_eval");
    if (rc == ZB_OK && red) rc = comm_publish(ctx, mb.seq, d == 1 ? 2 : (int)d + 1);
    if (rc == ZB_OK) { // the next round works on the new tables: put its kernel behind this one already
        Mle *mo[MAX_POLYS];
        for (uint32_t k = 0; k < d; k++) mo[k] = get_mle(ctx, out[k]);
        rc = prelaunch_next(ctx, out, d, mo, n / 2);
    }
    if (rc == ZB_OK) rc = wait_mail(ctx, mb.seq, mb.tagged ? (n == 2 ? (int)d : (d == 1 ? 2 : (int)d + 1)) : 0);
    if (rc) {
        tail_quiesce(ctx); // a pre-launched kernel may reference the tables that are dropped here
        for (uint32_t k = 0; k < d; k++) ctx->mles.erase(out[k]);
        return rc;
    }
    if (next) {
        if (n == 2)
            for (uint32_t k = 0; k < d; k++) next[k] = mail_word(ctx, (int)k);
        else
            evals_to_coeffs(d, ctx->h_mail, next);
    }
    return ZB_OK;
}

static int32_t fold_grid_impl(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint32_t nfold, const uint64_t *r, zb_mle *out,
                              uint64_t *grid) {
    tail_quiesce(ctx);
    Mle *ms[MAX_POLYS];
    int32_t rc = gather_polys(ctx, polys, d, ms);
    if (rc) return rc;
    if (nfold > 2 || !grid || (nfold && !r)) return ZB_ERR_BAD_ARGUMENT;
    const uint64_t n = ms[0]->n, m = n >> nfold;
    if ((m << nfold) != n || !fold_grid_ok(m)) return ZB_ERR_BAD_ARGUMENT;
    for (uint32_t t = 0; t < nfold; t++)
        if (r[t] >= bb::P) return ZB_ERR_NOT_CANONICAL;
    PolySet ps{};
    BufRef keep[MAX_POLYS];
    for (uint32_t k = 0; k < d; k++) {
        ps.src[k] = ms[k]->d();
        ps.dst[k] = ms[k]->d();
        keep[k] = ms[k]->buf;
    }
    if (out && nfold) {
        for (uint32_t k = 0; k < d; k++) {
            Mle *o = nullptr;
            rc = new_mle(ctx, m, &out[k], &o);
            if (rc) {
                for (uint32_t j = 0; j < k; j++) ctx->mles.erase(out[j]);
                return rc;
            }
            ps.dst[k] = o->d();
        }
    }
    const int np = d == 1 ? 2 : (int)d + 1, ns = np * np;
    const bool red = reduce_on_device(ctx);
    Mailbox mb = round_mailbox(ctx, red);
    static const char *names[3][3] = {{"grid_d1", "fold1_grid_d1", "fold2_grid_d1"},
                                      {"grid_d2", "fold1_grid_d2", "fold2_grid_d2"},
                                      {"grid_d3", "fold1_grid_d3", "fold2_grid_d3"}};
    // two-variable folds below 2^24 entries are latency-bound launches (>= 10 % of their duration is fixed cost): reported as a
    // family of their own, so that the bandwidth-bound launches' rate is not averaged with theirs
    static const char *small2[3] = {"fold2_grid_small_d1", "fold2_grid_small_d2", "fold2_grid_small_d3"};
    const char *scope = (nfold == 2 && n < (1ull << 24)) ? small2[d - 1] : names[d - 1][nfold];
    {
        ProfScope _ps(ctx, scope, (uint64_t)d * 4 * (n + (nfold ? m : 0)));
        launch_fold_grid((int)d, (int)nfold, ps, m, nfold ? (uint32_t)r[0] : 0, nfold > 1 ? (uint32_t)r[1] : 0, mb, ctx->sm_count,
                         ctx->stream);
    }
    rc = check_launch(ctx, "fold_grid");
    if (rc == ZB_OK && red) rc = comm_publish(ctx, mb.seq, ns);
    if (rc == ZB_OK) rc = wait_mail(ctx, mb.seq, mb.tagged ? ns : 0);
    if (rc) {
        if (out && nfold)
            for (uint32_t k = 0; k < d; k++) ctx->mles.erase(out[k]);
        return rc;
    }
    if (nfold && !out) {
        Mle *mm[MAX_POLYS];
        gather_polys(ctx, polys, d, mm); // the handle table may have been rehashed by new_mle: look the tables up again
        for (uint32_t k = 0; k < d; k++) mm[k]->n = m;
    }
    for (int k = 0; k < ns; k++) grid[k] = mail_word(ctx, k);
    return ZB_OK;
}

int32_t zb_prod_grid(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint64_t *grid) {
    return fold_grid_impl(ctx, polys, d, 0, nullptr, nullptr, grid);
}

int32_t zb_prod_fold_grid(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint32_t nfold, const uint64_t *r, zb_mle *out,
                          uint64_t *grid) {
    if (nfold < 1) return ZB_ERR_BAD_ARGUMENT;
    return fold_grid_impl(ctx, polys, d, nfold, r, out, grid);
}

int32_t zb_prod_fold_dump(zb_ctx *ctx, const zb_mle *polys, uint32_t d, uint32_t nfold, const uint64_t *r, uint32_t *tables) {
    if (d < 1 || d > (uint32_t)MAX_POLYS || !polys || !tables || nfold > 2 || (nfold && !r)) return ZB_ERR_BAD_ARGUMENT;
    tail_quiesce(ctx);
    Mle *ms[MAX_POLYS];
    int32_t rc = gather_polys(ctx, polys, d, ms);
    if (rc) return rc;
    const uint64_t n = ms[0]->n, m = n >> nfold;
    if (m < 1 || (m << nfold) != n || m > (1ull << PROD_DUMP_MAX_LOG2)) return ZB_ERR_BAD_ARGUMENT;
    for (uint32_t t = 0; t < nfold; t++)
        if (r[t] >= bb::P) return ZB_ERR_NOT_CANONICAL;
    PolySet ps{};
    for (uint32_t k = 0; k < d; k++) ps.src[k] = ps.dst[k] = ms[k]->d();
    Mailbox mb = ctx->mailbox();
    {
        ProfScope _ps(ctx, d == 1 ? "fold_dump_d1" : d == 2 ? "fold_dump_d2" : "fold_dump_d3", (uint64_t)d * 4 * (n + m));
        launch_fold_dump((int)d, (int)nfold, ps, m, nfold ? (uint32_t)r[0] : 0, nfold > 1 ? (uint32_t)r[1] : 0,
                         (uint32_t *)((uint8_t *)ctx->d_mail + DUMP_OFFSET), mb, ctx->stream);
    }
    rc = check_launch(ctx, "fold_dump");
    if (rc == ZB_OK) rc = wait_mail(ctx, mb.seq, 0); // a published table: fence + sequence number
    if (rc) return rc;
    memcpy(tables, (const uint8_t *)ctx->h_mail + DUMP_OFFSET, (size_t)d * m * sizeof(uint32_t));
    return ZB_OK;
}

int32_t zb_prod_collapse(zb_ctx *ctx, const zb_mle *polys, uint32_t d, const uint64_t *values) {
    if (d < 1 || d > (uint32_t)MAX_POLYS || !polys || !values) return ZB_ERR_BAD_ARGUMENT;
    tail_quiesce(ctx);
    Mle *ms[MAX_POLYS];
    int32_t rc = gather_polys(ctx, polys, d, ms);
    if (rc) return rc;
    uint32_t v[MAX_POLYS] = {0, 0, 0};
    PolySet ps{};
    for (uint32_t k = 0; k < d; k++) {
        if (values[k] >= bb::P) return ZB_ERR_NOT_CANONICAL;
        v[k] = (uint32_t)values[k];
        ps.src[k] = ps.dst[k] = ms[k]->d();
    }
    {
        ProfScope _ps(ctx, "fill", 4 * d);
        launch_fill_heads(ps, (int)d, v, ctx->stream);
    }
    LAUNCHED("fill_heads");
    for (uint32_t k = 0; k < d; k++) ms[k]->n = 1;
    return ZB_OK; // stream-ordered: every later call on the context sees the collapsed tables
}

/* ------------------------------------------------------------------ d = 1: several rounds per pass (linearity) */

int32_t zb_mle_block_sums(zb_ctx *ctx, zb_mle h, uint32_t k, uint64_t *sums) {
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (!sums || k < 1 || k > (uint32_t)LIN_WIDE_MAX_K || m->n < (4ull << k)) return ZB_ERR_BAD_ARGUMENT;
    if (k > (uint32_t)LIN_MAX_K) { // more than 32 blocks: one CTA per block, one self-validating word each (small tables)
        const Mailbox wmb = ctx->mailbox(); // takes a sequence number
        unsigned long long *d_words = (unsigned long long *)((uint8_t *)ctx->d_mail + DUMP_OFFSET);
        {
            ProfScope _ps(ctx, "block_sums_d1", m->n * 4);
            launch_block_sums_wide(m->d(), m->n, (int)k, d_words, wmb.seq, ctx->stream);
        }
        LAUNCHED("block_sums_wide");
        const volatile unsigned long long *w = (const volatile unsigned long long *)((uint8_t *)ctx->h_mail + DUMP_OFFSET);
        const int32_t rcw = wait_tagged_words(ctx, w, wmb.seq, 1ull << k);
        if (rcw) return rcw;
        for (uint32_t b = 0; b < (1u << k); b++) sums[b] = w[b] & 0xffffffffull;
        return ZB_OK;
    }
    Mailbox mb = ctx->mailbox();
    {
        ProfScope _ps(ctx, "block_sums_d1", m->n * 4);
        launch_block_sums(m->d(), m->n, (int)k, mb, ctx->sm_count, ctx->stream);
    }
    LAUNCHED("block_sums");
    int32_t rc = wait_mail(ctx, mb.seq, 1 << k);
    if (rc) return rc;
    for (uint32_t b = 0; b < (1u << k); b++) sums[b] = mail_word(ctx, (int)b);
    return ZB_OK;
}

int32_t zb_mle_fold_multi(zb_ctx *ctx, zb_mle h, uint32_t k_fold, const uint64_t *r, zb_mle *out, uint32_t k_next, uint64_t *sums) {
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (!r || !sums || k_fold < 1 || k_fold > (uint32_t)LIN_WIDE_MAX_K) return ZB_ERR_BAD_ARGUMENT;
    const uint64_t n = m->n, mm = n >> k_fold;
    if (mm < 4 || (mm << k_fold) != n) return ZB_ERR_BAD_ARGUMENT;
    const bool dump = (1ull << k_next) == mm;
    if (k_fold > (uint32_t)LIN_MAX_K && !dump) return ZB_ERR_BAD_ARGUMENT; // more than 5 variables only when the folded table is published
    static const bool wide_on = [] {
        const char *e = getenv("ZB_LIN_WIDE"); // 0: published tables through the streaming kernel (k_fold <= 5 only)
        return !(e && *e == '0');
    }();
    if (dump && k_next <= (uint32_t)LIN_DUMP_MAX_LOG2 && (wide_on || k_fold > (uint32_t)LIN_MAX_K)) {
        // small table, published whole: the wide kernel (all rows of a column in one CTA, self-validating words, no ticket)
        for (uint32_t j = 0; j < k_fold; j++)
            if (r[j] >= bb::P) return ZB_ERR_NOT_CANONICAL;
        WideWeights ww{};
        ww.w[0] = bb::R_MOD_P;
        for (uint32_t j = 0; j < k_fold; j++) {
            const uint32_t rj = (uint32_t)r[j], nj = bb::sub(1u, rj);
            for (int t = (1 << j) - 1; t >= 0; t--) {
                const uint32_t base = ww.w[t];
                ww.w[2 * t] = bb::mul(base, nj);
                ww.w[2 * t + 1] = bb::mul(base, rj);
            }
        }
        const uint32_t *wsrc = m->d();
        uint32_t *wdst = m->d();
        BufRef wkeep = m->buf;
        if (out) {
            Mle *o = nullptr;
            int32_t rc0 = new_mle(ctx, mm, out, &o);
            if (rc0) return rc0;
            wdst = o->d();
        }
        const Mailbox wmb = ctx->mailbox();
        unsigned long long *d_words = (unsigned long long *)((uint8_t *)ctx->d_mail + DUMP_OFFSET);
        {
            ProfScope _ps(ctx, "foldk_d1", (n + mm) * 4);
            launch_foldk_wide(wsrc, wdst, n, (int)k_fold, ww, d_words, wmb.seq, ctx->stream);
        }
        int32_t rcw = check_launch(ctx, "foldk_wide");
        const volatile unsigned long long *w = (const volatile unsigned long long *)((uint8_t *)ctx->h_mail + DUMP_OFFSET);
        if (rcw == ZB_OK) rcw = wait_tagged_words(ctx, w, wmb.seq, mm);
        if (rcw) {
            if (out) {
                ctx->mles.erase(*out);
                *out = 0;
            }
            return rcw;
        }
        if (!out) get_mle(ctx, h)->n = mm;
        for (uint64_t i = 0; i < mm; i++) sums[i] = w[i] & 0xffffffffull;
        return ZB_OK;
    }
    if (dump ? k_next > (uint32_t)LIN_DUMP_MAX_LOG2 : (k_next < 1 || k_next > (uint32_t)LIN_MAX_K || mm < (4ull << k_next)))
        return ZB_ERR_BAD_ARGUMENT;
    for (uint32_t j = 0; j < k_fold; j++)
        if (r[j] >= bb::P) return ZB_ERR_NOT_CANONICAL;
    // eq weights, index = the k_fold top index bits with r[0] <-> the most significant one, in Montgomery form
    FoldWeights fw{};
    fw.w[0] = bb::R_MOD_P; // 1 in Montgomery form
    for (uint32_t j = 0; j < k_fold; j++) { // after step j: w[0 .. 2^(j+1)) for the bits b_0..b_j (b_0 = MSB so far)
        const uint32_t rj = (uint32_t)r[j], nj = bb::sub(1u, rj);
        for (int t = (1 << j) - 1; t >= 0; t--) {
            const uint32_t base = fw.w[t];
            fw.w[2 * t] = bb::mul(base, nj);
            fw.w[2 * t + 1] = bb::mul(base, rj);
        }
    }
    const uint32_t *src = m->d();
    uint32_t *dst = m->d();
    BufRef keep = m->buf;
    if (out) {
        Mle *o = nullptr;
        int32_t rc = new_mle(ctx, mm, out, &o);
        if (rc) return rc;
        dst = o->d();
    }
    Mailbox mb = ctx->mailbox();
    unsigned long long *d_dump = dump ? (unsigned long long *)((uint8_t *)ctx->d_mail + DUMP_OFFSET) : nullptr;
    {
        ProfScope _ps(ctx, "foldk_d1", (n + mm) * 4);
        launch_foldk_sums(src, dst, n, (int)k_fold, fw, (int)k_next, d_dump, mb, ctx->sm_count, ctx->stream);
    }
    int32_t rc = check_launch(ctx, "foldk_sums");
    if (rc == ZB_OK) rc = wait_mail(ctx, mb.seq, dump ? 0 : 1 << k_next); // a published table: fence + sequence number
    if (rc) {
        if (out) {
            ctx->mles.erase(*out);
            *out = 0;
        }
        return rc;
    }
    if (!out) get_mle(ctx, h)->n = mm;
    if (dump) {
        const uint32_t *t = (const uint32_t *)((uint8_t *)ctx->h_mail + DUMP_OFFSET);
        for (uint64_t i = 0; i < mm; i++) sums[i] = t[i];
    } else
        for (uint32_t b = 0; b < (1u << k_next); b++) sums[b] = mail_word(ctx, (int)b);
    return ZB_OK;
}

int32_t zb_mle_collapse(zb_ctx *ctx, zb_mle h, uint64_t value) {
    tail_quiesce(ctx);
    Mle *m = get_mle(ctx, h);
    if (!m) return ZB_ERR_BAD_HANDLE;
    if (value >= bb::P) return ZB_ERR_NOT_CANONICAL;
    {
        ProfScope _ps(ctx, "fill", 4);
        launch_fill(m->d(), 1, (uint32_t)value, ctx->stream);
    }
    LAUNCHED("fill");
    m->n = 1;
    return zb_sync(ctx);
}

/* ------------------------------------------------------------------ Merkle */

static int32_t build_batch(zb_ctx *ctx, Tree **ts, uint32_t count, uint8_t *roots, bool with_leaves = true) {
    const uint64_t padded = ts[0]->padded;
    const uint32_t height = ts[0]->height;
    MerkleBatch b{};
    b.count = count;
    for (uint32_t t = 0; t < count; t++) {
        b.values[t] = (const uint32_t *)ts[t]->values->ptr;
        b.n_values[t] = ts[t]->n_values;
        b.tree[t] = (uint8_t *)ts[t]->store->ptr;
    }
    if (with_leaves) {
        {
            ProfScope _ps(ctx, "merkle_leaves", padded * 36 * count);
            launch_merkle_leaves(b, padded, ctx->stream);
        }
        LAUNCHED("merkle_leaves");
    }
    uint32_t level = 0;
    while ((padded >> level) > MERKLE_TOP_WIDTH) {
        {
            ProfScope _ps(ctx, "merkle_level", (padded >> (level + 1)) * 96 * count);
            launch_merkle_level(b, padded, level, ctx->stream);
        }
        LAUNCHED("merkle_level");
        level++;
    }
    if (level < height) {
        {
            ProfScope _ps(ctx, "merkle_top", (padded >> level) * 96 * count);
            launch_merkle_top(b, padded, level, ctx->stream);
        }
        LAUNCHED("merkle_top");
    }
    Mailbox mb = ctx->mailbox();
    {
        ProfScope _ps(ctx, "merkle_roots", 64ull * count);
        launch_merkle_roots(b, padded, ctx->d_bulk(), mb, ctx->stream);
    }
    LAUNCHED("merkle_roots");
    int32_t rc = wait_mail(ctx, mb.seq);
    if (rc) return rc;
    for (uint32_t t = 0; t < count; t++) {
        memcpy(ts[t]->root, ctx->h_bulk() + 32 * t, 32);
        if (roots) memcpy(roots + 32 * t, ts[t]->root, 32);
    }
    return ZB_OK;
}

static int32_t new_tree(zb_ctx *ctx, BufRef values, uint64_t n_values, zb_tree *out, Tree **tp) {
    uint64_t padded = 1;
    while (padded < n_values) padded <<= 1;
    BufRef store;
    int32_t rc = dev_alloc(ctx, (2 * padded - 1) * 32, &store);
    if (rc) return rc;
    uint64_t h = ctx->next_handle++;
    Tree t{};
    t.values = values;
    t.n_values = n_values;
    t.padded = padded;
    t.height = (uint32_t)__builtin_ctzll(padded);
    t.store = store;
    ctx->trees[h] = t;
    *out = h;
    *tp = &ctx->trees[h];
    return ZB_OK;
}

int32_t zb_merkle_build(zb_ctx *ctx, const zb_mle *polys, uint32_t count, zb_tree *trees, uint8_t *roots) {
    if (group_front(ctx) && polys && count && is_group_handle(polys[0])) return zg_merkle_build(ctx, polys, count, trees, roots);
    tail_quiesce(ctx);
    if (!polys || !trees || count == 0) return ZB_ERR_BAD_ARGUMENT;
    uint64_t n0 = 0;
    for (uint32_t i = 0; i < count; i++) {
        Mle *m = get_mle(ctx, polys[i]);
        if (!m) return ZB_ERR_BAD_HANDLE;
        if (i == 0) n0 = m->n;
        else if (m->n != n0) return ZB_ERR_DIFFERENT_NUM_VARS;
    }
    // a failure in a later batch must not leave the trees of the completed batches behind
    auto undo = [&](uint32_t done, int32_t rc) {
        cudaStreamSynchronize(ctx->stream);
        for (uint32_t j = 0; j < done; j++) {
            ctx->trees.erase(trees[j]);
            trees[j] = 0;
        }
        return rc;
    };
    for (uint32_t base = 0; base < count; base += MAX_BATCH) {
        uint32_t c = count - base < (uint32_t)MAX_BATCH ? count - base : MAX_BATCH;
        std::vector<zb_tree> hs(c);
        for (uint32_t t = 0; t < c; t++) {
            Mle *m = get_mle(ctx, polys[base + t]);
            // the tree keeps its OWN copy of the values, like SimpleMerkleTree.build (merkle_tree.zig:291): the polynomial
            // may be folded in place afterwards (consuming sumcheck) without changing what `open` reports
            BufRef vals;
            int32_t rc = dev_alloc(ctx, m->n * sizeof(uint32_t), &vals);
            cudaError_t ce = cudaSuccess;
            if (rc == ZB_OK)
                ce = cudaMemcpyAsync(vals->ptr, m->d(), m->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
            if (rc == ZB_OK && ce != cudaSuccess) rc = cuda_fail(ctx, ce, "cudaMemcpyAsync(tree values)");
            Tree *tp = nullptr;
            if (rc == ZB_OK) rc = new_tree(ctx, vals, m->n, &hs[t], &tp);
            if (rc) {
                cudaStreamSynchronize(ctx->stream);
                for (uint32_t j = 0; j < t; j++) ctx->trees.erase(hs[j]);
                return undo(base, rc);
            }
        }
        Tree *ts[MAX_BATCH];
        for (uint32_t t = 0; t < c; t++) ts[t] = get_tree(ctx, hs[t]); // map is stable now
        int32_t rc = build_batch(ctx, ts, c, roots ? roots + 32 * (size_t)base : nullptr);
        if (rc) {
            cudaStreamSynchronize(ctx->stream);
            for (uint32_t t = 0; t < c; t++) ctx->trees.erase(hs[t]);
            return undo(base, rc);
        }
        for (uint32_t t = 0; t < c; t++) trees[base + t] = hs[t];
    }
    return ZB_OK;
}

int32_t zb_merkle_build_values(zb_ctx *ctx, const uint64_t *values, uint64_t n, zb_tree *tree, uint8_t root[32]) {
    tail_quiesce(ctx);
    if (n == 0) return ZB_ERR_EMPTY_VALUES; // merkle_tree.zig:284
    if (!values || !tree) return ZB_ERR_BAD_ARGUMENT;
    BufRef vals;
    int32_t rc = dev_alloc(ctx, n * sizeof(uint32_t), &vals);
    if (rc) return rc;
    rc = upload_narrow(ctx, values, n, (uint32_t *)vals->ptr);
    if (rc) return rc;
    Tree *tp = nullptr;
    rc = new_tree(ctx, vals, n, tree, &tp);
    if (rc) return rc;
    rc = build_batch(ctx, &tp, 1, root);
    if (rc) {
        ctx->trees.erase(*tree);
        *tree = 0;
    }
    return rc;
}

int32_t zb_merkle_info(zb_ctx *ctx, zb_tree h, uint64_t *n_values, uint32_t *height, uint8_t root[32]) {
    if (group_front(ctx) && is_group_handle(h)) {
        GTree *gt = get_gtree(ctx, h);
        if (!gt) return ZB_ERR_BAD_HANDLE;
        if (n_values) *n_values = gt->n_values;
        if (height) *height = gt->height;
        if (root) memcpy(root, gt->root, 32);
        return ZB_OK;
    }
    Tree *t = get_tree(ctx, h);
    if (!t) return ZB_ERR_BAD_HANDLE;
    if (n_values) *n_values = t->n_values;
    if (height) *height = t->height;
    if (root) memcpy(root, t->root, 32);
    return ZB_OK;
}

int32_t zb_merkle_open(zb_ctx *ctx, zb_tree h, uint64_t index, uint8_t *siblings, uint8_t *dirs, uint64_t *leaf_value) {
    if (group_front(ctx) && is_group_handle(h)) return zg_merkle_open(ctx, h, index, siblings, dirs, leaf_value);
    tail_quiesce(ctx);
    Tree *t = get_tree(ctx, h);
    if (!t) return ZB_ERR_BAD_HANDLE;
    if (index >= t->n_values) return ZB_ERR_INDEX_OUT_OF_BOUNDS; // merkle_tree.zig:325
    if (t->height > 64) return ZB_ERR_BAD_ARGUMENT;
    // leaf value rides along in the bulk area after the digests: copy it with the same gather launch's stream
    Mailbox mb = ctx->mailbox();
    CK(cudaMemcpyAsync(ctx->h_bulk() + 64 * 32, (const uint32_t *)t->values->ptr + index, sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, ctx->stream));
    {
        ProfScope _ps(ctx, "merkle_path", 64ull * t->height);
        launch_merkle_path((const uint8_t *)t->store->ptr, t->padded, t->height, index, ctx->d_bulk(), mb, ctx->stream);
    }
    LAUNCHED("merkle_path");
    int32_t rc = wait_mail(ctx, mb.seq);
    if (rc) return rc;
    if (t->height) memcpy(siblings, ctx->h_bulk(), (size_t)t->height * 32);
    for (uint32_t l = 0; l < t->height; l++) dirs[l] = (uint8_t)((index >> l) & 1); // :341-345
    if (leaf_value) {
        uint32_t v;
        memcpy(&v, ctx->h_bulk() + 64 * 32, sizeof(v));
        *leaf_value = v;
    }
    return ZB_OK;
}

// open for `count` trees of equal shape, one leaf each, in one launch and one read-back
int32_t zb_merkle_open_batch(zb_ctx *ctx, const zb_tree *trees, uint32_t count, const uint64_t *indices, uint8_t *siblings,
                             uint8_t *dirs, uint64_t *leaf_values) {
    tail_quiesce(ctx);
    if (!trees || !indices || count == 0 || count > 65535) return ZB_ERR_BAD_ARGUMENT;
    uint64_t padded = 0;
    uint32_t height = 0;
    std::vector<uint64_t> blob(3 * (size_t)count);
    for (uint32_t i = 0; i < count; i++) {
        Tree *t = get_tree(ctx, trees[i]);
        if (!t) return ZB_ERR_BAD_HANDLE;
        if (indices[i] >= t->n_values) return ZB_ERR_INDEX_OUT_OF_BOUNDS; // merkle_tree.zig:325
        if (i == 0) padded = t->padded, height = t->height;
        else if (t->padded != padded) return ZB_ERR_DIFFERENT_NUM_VARS;
        blob[i] = (uint64_t)(uintptr_t)t->store->ptr;
        blob[count + i] = (uint64_t)(uintptr_t)t->values->ptr;
        blob[2 * (size_t)count + i] = indices[i];
    }
    if (height > 64) return ZB_ERR_BAD_ARGUMENT;
    const size_t path_bytes = (size_t)count * height * 32, out_bytes = path_bytes + (size_t)count * 4, blob_bytes = blob.size() * 8;
    uint8_t *h = nullptr;
    int32_t rc = zb_host_scratch(ctx, blob_bytes + out_bytes, (void **)&h);
    if (rc) return rc;
    memcpy(h, blob.data(), blob_bytes);
    BufRef d_blob, d_out;
    rc = dev_alloc(ctx, blob_bytes, &d_blob);
    if (rc == ZB_OK) rc = dev_alloc(ctx, out_bytes + 64, &d_out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d_blob->ptr, h, blob_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t *db = (const uint64_t *)d_blob->ptr;
    {
        ProfScope _ps(ctx, "merkle_path_batch", 64ull * height * count);
        launch_merkle_path_batch((const uint8_t *const *)db, (const uint32_t *const *)(db + count), db + 2 * (size_t)count, count, padded,
                                 height, (uint8_t *)d_out->ptr, (uint32_t *)((uint8_t *)d_out->ptr + path_bytes), ctx->stream);
    }
    LAUNCHED("merkle_path_batch");
    CK(cudaMemcpyAsync(h + blob_bytes, d_out->ptr, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (siblings && height) memcpy(siblings, h + blob_bytes, path_bytes);
    for (uint32_t i = 0; i < count; i++) {
        if (dirs)
            for (uint32_t l = 0; l < height; l++) dirs[(size_t)i * height + l] = (uint8_t)((indices[i] >> l) & 1); // :341-345
        if (leaf_values) {
            uint32_t val;
            memcpy(&val, h + blob_bytes + path_bytes + 4 * (size_t)i, 4);
            leaf_values[i] = val;
        }
    }
    return ZB_OK;
}

int32_t zb_merkle_leaf_hashes(zb_ctx *ctx, zb_tree h, uint8_t *out, uint64_t n_digests) {
    tail_quiesce(ctx);
    Tree *t = get_tree(ctx, h);
    if (!t) return ZB_ERR_BAD_HANDLE;
    if (n_digests > 2 * t->padded - 1) return ZB_ERR_BAD_ARGUMENT;
    CK(cudaMemcpyAsync(out, t->store->ptr, n_digests * 32, cudaMemcpyDeviceToHost, ctx->stream));
    return zb_sync(ctx);
}

int32_t zb_merkle_free(zb_ctx *ctx, zb_tree h) {
    if (group_front(ctx) && is_group_handle(h)) return zg_merkle_free(ctx, h);
    tail_quiesce(ctx);
    if (!ctx->trees.erase(h)) return ZB_ERR_BAD_HANDLE;
    return ZB_OK;
}

/* ------------------------------------------------------------------ Lasso */

int32_t zb_xxh3_rows(zb_ctx *ctx, const uint64_t *rows, uint64_t n_rows, uint32_t arity, uint64_t n_padded, zb_mle *out) {
    tail_quiesce(ctx);
    int32_t rc = check_pow2(n_padded);
    if (rc) return rc;
    if (n_rows > n_padded || arity == 0 || (!rows && n_rows) || !out) return ZB_ERR_BAD_ARGUMENT;
    Mle *m = nullptr;
    rc = new_mle(ctx, n_padded, out, &m);
    if (rc) return rc;
    // the rows are field elements (canonical, < 2^31): narrow them on the way up like any other table, then hash
    if (n_rows) {
        BufRef drows;
        rc = dev_alloc(ctx, n_rows * arity * sizeof(uint32_t), &drows);
        if (rc == ZB_OK) rc = upload_narrow(ctx, rows, n_rows * arity, (uint32_t *)drows->ptr);
        if (rc) {
            ctx->mles.erase(*out);
            *out = 0;
            return rc;
        }
        {
            ProfScope _ps(ctx, "xxh3_rows", n_rows * (4ull * arity) + n_padded * 4);
            launch_xxh3_rows((const uint32_t *)drows->ptr, n_rows, arity, n_padded, m->d(), ctx->stream);
        }
        LAUNCHED("xxh3_rows");
    } else {
        {
            ProfScope _ps(ctx, "fill", n_padded * 4);
            launch_fill(m->d(), n_padded, 0, ctx->stream);
        }
        LAUNCHED("fill");
    }
    return zb_sync(ctx);
}

int32_t zb_xxh3_rows_stream(zb_ctx *ctx, const uint64_t *rows, uint64_t n_rows, uint32_t arity, uint64_t n_padded, zb_mle *out,
                            uint32_t *host_mirror, uint64_t *avail) {
    tail_quiesce(ctx);
    int32_t rc = check_pow2(n_padded);
    if (rc) return rc;
    if (n_rows > n_padded || arity == 0 || (!rows && n_rows) || !out || !host_mirror || !avail) return ZB_ERR_BAD_ARGUMENT;
    Mle *m = nullptr;
    rc = new_mle(ctx, n_padded, out, &m);
    if (rc) return rc;
    auto fail = [&](int32_t e) {
        cudaStreamSynchronize(ctx->stream);
        ctx->mles.erase(*out);
        *out = 0;
        return e;
    };
    const uint64_t CH = 1ull << ctx->lasso_chunk_log2; // rows per streamed chunk (option "lasso_chunk_log2")
    if (n_rows) {
        BufRef drows;
        rc = dev_alloc(ctx, n_rows * arity * sizeof(uint32_t), &drows);
        if (rc) return fail(rc);
        uint32_t *d_rows = (uint32_t *)drows->ptr;
        for (uint64_t r0 = 0; r0 < n_rows; r0 += CH) {
            const uint64_t r1 = r0 + CH < n_rows ? r0 + CH : n_rows;
            // drains the stream: the previous chunk's hashes have landed in the mirror when this returns
            rc = upload_narrow(ctx, rows + r0 * arity, (r1 - r0) * arity, d_rows + r0 * arity);
            if (rc) return fail(rc);
            __atomic_store_n(avail, r0, __ATOMIC_RELEASE);
            const uint64_t n_out = r1 == n_rows ? n_padded - r0 : r1 - r0; // the last chunk also writes the zero padding
            {
                ProfScope _ps(ctx, "xxh3_rows", (r1 - r0) * (4ull * arity) + n_out * 4);
                launch_xxh3_rows(d_rows + r0 * arity, r1 - r0, arity, n_out, m->d() + r0, ctx->stream);
            }
            rc = check_launch(ctx, "xxh3_rows");
            if (rc) return fail(rc);
            const cudaError_t ce = cudaMemcpyAsync(host_mirror + r0, m->d() + r0, n_out * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (ce != cudaSuccess) return fail(cuda_fail(ctx, ce, "cudaMemcpyAsync(mirror)"));
        }
        rc = zb_sync(ctx); // before drows goes back to the pool
        if (rc) return fail(rc);
    } else {
        {
            ProfScope _ps(ctx, "fill", n_padded * 4);
            launch_fill(m->d(), n_padded, 0, ctx->stream);
        }
        rc = check_launch(ctx, "fill");
        if (rc) return fail(rc);
        memset(host_mirror, 0, n_padded * sizeof(uint32_t));
        rc = zb_sync(ctx);
        if (rc) return fail(rc);
    }
    __atomic_store_n(avail, n_padded, __ATOMIC_RELEASE);
    return ZB_OK;
}

int32_t zb_table_mle(zb_ctx *ctx, int32_t op, uint32_t bits, zb_mle *out) {
    tail_quiesce(ctx);
    if (op < 0 || op > 2 || bits == 0 || bits > 15 || !out) return ZB_ERR_BAD_ARGUMENT;
    Mle *m = nullptr;
    int32_t rc = new_mle(ctx, 1ull << (2 * bits), out, &m);
    if (rc) return rc;
    {
        ProfScope _ps(ctx, "table_mle", 4ull << (2 * bits));
        launch_table_mle(op, bits, m->d(), ctx->stream);
    }
    LAUNCHED("table_mle");
    return zb_sync(ctx);
}

/* ------------------------------------------------------------------ witness packing */

int32_t zb_witness_pack(zb_ctx *ctx, const uint64_t *cols, uint64_t num_steps, uint32_t n_cols, uint32_t n_hold, zb_mle *out,
                        uint32_t *num_vars) {
    tail_quiesce(ctx);
    if (n_cols == 0 || n_cols > 64 || n_hold > n_cols || !out || (!cols && num_steps)) return ZB_ERR_BAD_ARGUMENT;
    uint32_t v = 0;
    while ((1ull << v) < num_steps) v++; // witness.zig:36-39 (0 steps -> one zero entry)
    const uint64_t padded = 1ull << v;
    if (num_vars) *num_vars = v;
    uint32_t *ptrs[64];
    for (uint32_t c = 0; c < n_cols; c++) {
        Mle *m = nullptr;
        int32_t rc = new_mle(ctx, padded, &out[c], &m);
        if (rc) {
            for (uint32_t j = 0; j < c; j++) ctx->mles.erase(out[j]);
            return rc;
        }
        ptrs[c] = m->d();
    }
    auto fail = [&](int32_t rc) {
        for (uint32_t c = 0; c < n_cols; c++) ctx->mles.erase(out[c]);
        return rc;
    };
    // padding values: F.init(last real value) of the "hold" columns (witness.zig:80-87, :116-123)
    uint32_t h_last[64] = {0};
    for (uint32_t c = 0; c < n_hold && num_steps; c++) h_last[c] = (uint32_t)(cols[(size_t)c * num_steps + num_steps - 1] % bb::P);
    BufRef d_last, stage;
    int32_t rc = dev_alloc(ctx, sizeof(h_last), &d_last);
    if (rc) return fail(rc);
    CK(cudaMemcpyAsync(d_last->ptr, h_last, sizeof(h_last), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // h_last is a stack buffer
    uint64_t chunk = STAGE_ELEMS / n_cols;
    if (chunk > padded) chunk = padded;
    rc = dev_alloc(ctx, chunk * n_cols * sizeof(uint64_t), &stage);
    if (rc) return fail(rc);
    for (uint64_t step0 = 0; step0 < padded; step0 += chunk) {
        const uint64_t real = step0 < num_steps ? (num_steps - step0 < chunk ? num_steps - step0 : chunk) : 0;
        for (uint32_t c = 0; c < n_cols && real; c++) {
            rc = staged_h2d(ctx, (uint64_t *)stage->ptr + (size_t)c * chunk, cols + (size_t)c * num_steps + step0, real * sizeof(uint64_t));
            if (rc) return fail(rc);
        }
        {
            ProfScope _ps(ctx, "witness_pack", (uint64_t)n_cols * (real * 8 + chunk * 4));
            launch_witness_pack((const uint64_t *)stage->ptr, chunk, step0, num_steps, padded, n_cols, n_hold,
                                (const uint32_t *)d_last->ptr, ptrs, ctx->stream);
        }
        rc = check_launch(ctx, "witness_pack");
        if (rc) return fail(rc);
    }
    rc = zb_sync(ctx);
    return rc ? fail(rc) : ZB_OK;
}

// WitnessGenerator.generate + the commit phase of Prover.generateCommitments (prover.zig:405-410) as ONE pipeline: column c
// goes up on a copy stream while the SMs pack and leaf-hash column c-1, so the 8-byte-per-value trace upload (PCIe) and the
// leaf hashing (integer pipe) overlap instead of adding up; the upper levels of all trees follow as one batched build.
int32_t zb_witness_pack_commit(zb_ctx *ctx, const uint64_t *cols, uint64_t num_steps, uint32_t n_cols, uint32_t n_hold, zb_mle *out,
                               uint32_t *num_vars, zb_tree *trees, uint8_t *roots) {
    tail_quiesce(ctx);
    if (n_cols == 0 || n_cols > (uint32_t)MAX_BATCH || n_hold > n_cols || !out || !trees || !cols || num_steps == 0)
        return ZB_ERR_BAD_ARGUMENT;
    uint32_t v = 0;
    while ((1ull << v) < num_steps) v++;
    const uint64_t padded = 1ull << v;
    if (num_vars) *num_vars = v;
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 3; b++) {
            CK(cudaEventCreateWithFlags(&ctx->pipe_up[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->pipe_used[b], cudaEventDisableTiming));
        }
    }
    uint32_t made = 0, made_trees = 0;
    auto fail = [&](int32_t rc) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        for (uint32_t c = 0; c < made; c++) ctx->mles.erase(out[c]);
        for (uint32_t c = 0; c < made_trees; c++) ctx->trees.erase(trees[c]);
        return rc;
    };
    uint32_t *ptrs[MAX_BATCH];
    for (uint32_t c = 0; c < n_cols; c++, made++) {
        Mle *m = nullptr;
        const int32_t rc = new_mle(ctx, padded, &out[c], &m);
        if (rc) return fail(rc);
        ptrs[c] = m->d();
    }
    Tree *ts[MAX_BATCH];
    for (uint32_t c = 0; c < n_cols; c++, made_trees++) {
        BufRef vals; // the tree's own copy of the values (merkle_tree.zig:291)
        int32_t rc = dev_alloc(ctx, padded * sizeof(uint32_t), &vals);
        Tree *tp = nullptr;
        if (rc == ZB_OK) rc = new_tree(ctx, vals, padded, &trees[c], &tp);
        if (rc) return fail(rc);
    }
    for (uint32_t c = 0; c < n_cols; c++) ts[c] = get_tree(ctx, trees[c]); // the table no longer grows
    uint32_t h_last[MAX_BATCH] = {0};
    for (uint32_t c = 0; c < n_hold; c++) h_last[c] = (uint32_t)(cols[(size_t)c * num_steps + num_steps - 1] % bb::P);
    BufRef d_last, slots;
    int32_t rc = dev_alloc(ctx, sizeof(h_last), &d_last);
    if (rc == ZB_OK) rc = dev_alloc(ctx, 3 * num_steps * sizeof(uint64_t), &slots);
    if (rc) return fail(rc);
    cudaError_t ce = cudaMemcpyAsync(d_last->ptr, h_last, sizeof(h_last), cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream); // h_last is a stack buffer
    if (ce != cudaSuccess) return fail(cuda_fail(ctx, ce, "witness_pack_commit: padding values"));
    for (uint32_t c = 0; c < n_cols; c++) {
        const int slot = (int)(c % 3);
        uint64_t *stage = (uint64_t *)slots->ptr + (size_t)slot * num_steps;
        if (c >= 3) ce = cudaStreamWaitEvent(ctx->copy_stream, ctx->pipe_used[slot], 0); // the pack kernel has drained the slot
        if (ce != cudaSuccess) return fail(cuda_fail(ctx, ce, "cudaStreamWaitEvent"));
        rc = staged_h2d(ctx, stage, cols + (size_t)c * num_steps, num_steps * sizeof(uint64_t), ctx->copy_stream);
        if (rc) return fail(rc);
        ce = cudaEventRecord(ctx->pipe_up[slot], ctx->copy_stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stream, ctx->pipe_up[slot], 0);
        if (ce != cudaSuccess) return fail(cuda_fail(ctx, ce, "cudaEventRecord"));
        {
            ProfScope _ps(ctx, "witness_pack", num_steps * 8 + padded * 4);
            launch_witness_pack(stage, padded, 0, num_steps, padded, 1, c < n_hold ? 1u : 0u, (const uint32_t *)d_last->ptr + c, &ptrs[c],
                                ctx->stream);
        }
        rc = check_launch(ctx, "witness_pack");
        if (rc) return fail(rc);
        ce = cudaEventRecord(ctx->pipe_used[slot], ctx->stream);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(ts[c]->values->ptr, ptrs[c], padded * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce != cudaSuccess) return fail(cuda_fail(ctx, ce, "witness_pack_commit: tree values"));
        MerkleBatch one{};
        one.count = 1;
        one.values[0] = (const uint32_t *)ts[c]->values->ptr;
        one.n_values[0] = padded;
        one.tree[0] = (uint8_t *)ts[c]->store->ptr;
        {
            ProfScope _ps(ctx, "merkle_leaves", padded * 36);
            launch_merkle_leaves(one, padded, ctx->stream);
        }
        rc = check_launch(ctx, "merkle_leaves");
        if (rc) return fail(rc);
    }
    rc = build_batch(ctx, ts, n_cols, roots, /*with_leaves=*/false);
    if (rc) return fail(rc);
    ce = cudaStreamSynchronize(ctx->copy_stream);
    if (ce != cudaSuccess) return fail(cuda_fail(ctx, ce, "copy stream"));
    return ZB_OK;
}

/* ------------------------------------------------------------------ multi-GPU (NCCL, dlopen'ed) */

struct NcclId { // ncclUniqueId (nccl.h:38-39), passed by value
    char internal[128];
};
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int NCCL_UINT64 = 5, NCCL_UINT32 = 3, NCCL_SUM = 0; // nccl.h: ncclUint64, ncclUint32, ncclSum

int32_t nccl_load(zb_ctx *ctx, const char *path) {
    if (g_nccl.lib) return ZB_OK;
    const char *cands[3] = {path, getenv("ZIGZ_NCCL_LIB"), "libnccl.so.2"};
    void *h = nullptr;
    for (const char *c : cands)
        if (c && *c && (h = dlopen(c, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) {
        if (ctx) ctx->last_error = std::string("dlopen libnccl: ") + dlerror();
        return ZB_ERR_NCCL;
    }
    NcclApi a;
    a.lib = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
        if (ctx) ctx->last_error = "libnccl lacks a required symbol";
        return ZB_ERR_NCCL;
    }
    g_nccl = a;
    return ZB_OK;
}

int32_t nccl_fail(zb_ctx *ctx, int r, const char *what) {
    if (ctx) ctx->last_error = std::string(what) + ": " + g_nccl.GetErrorString(r);
    return ZB_ERR_NCCL;
}
} // namespace

extern "C++" {
namespace {
// sum the first nwords of the device exchange buffer over all ranks (stream order, behind the kernel that wrote
// them) and publish them mod p with sequence number `seq`
int32_t comm_publish(zb_ctx *ctx, unsigned long long seq, int nwords) {
    if (reduce_p2p(ctx)) return ZB_OK; // already summed over NVLink inside the kernel
    int r = g_nccl.AllReduce(ctx->d_comm, ctx->d_comm, (size_t)nwords, NCCL_UINT64, NCCL_SUM, ctx->nccl_comm, ctx->stream);
    if (r) return nccl_fail(ctx, r, "ncclAllReduce");
    launch_publish_reduced(ctx->d_comm, nwords, ctx->d_mail, seq, ctx->stream);
    return check_launch(ctx, "publish_reduced");
}
} // namespace
} // extern "C++"

int32_t zb_comm_allgather_cyclic_batch(zb_ctx *ctx, const zb_mle *locals, uint32_t count, zb_mle *outs) {
    tail_quiesce(ctx);
    if (!locals || !outs || count < 1 || count > (uint32_t)MAX_POLYS) return ZB_ERR_BAD_ARGUMENT;
    uint64_t n = 0;
    BufRef src[MAX_POLYS];
    for (uint32_t k = 0; k < count; k++) {
        Mle *m = get_mle(ctx, locals[k]);
        if (!m) return ZB_ERR_BAD_HANDLE;
        if (k && m->n != n) return ZB_ERR_DIFFERENT_NUM_VARS;
        n = m->n;
        src[k] = m->buf;
        outs[k] = 0;
    }
    auto drop = [&]() {
        for (uint32_t k = 0; k < count; k++)
            if (outs[k]) {
                ctx->mles.erase(outs[k]);
                outs[k] = 0;
            }
    };
    if (ctx->world == 1) {
        int32_t rc = ZB_OK;
        for (uint32_t k = 0; k < count && rc == ZB_OK; k++) rc = zb_mle_clone(ctx, locals[k], &outs[k]);
        if (rc) drop();
        return rc;
    }
    uint32_t *dst[MAX_POLYS] = {nullptr, nullptr, nullptr};
    int32_t rc = ZB_OK;
    for (uint32_t k = 0; k < count && rc == ZB_OK; k++) {
        Mle *o = nullptr;
        rc = new_mle(ctx, n * ctx->world, &outs[k], &o);
        if (rc == ZB_OK) dst[k] = o->d();
    }
    static const bool xchg_gather = [] {
        const char *e = getenv("ZB_GATHER_XCHG"); // 0: NCCL / host-rendezvous all-gather (the fallback paths below)
        return !(e && *e == '0');
    }();
    if (rc == ZB_OK && xchg_gather && ctx->d_xchg_view && n <= (1ull << XCHG_GATHER_MAX_LOG2)) {
        // peers attached (CUDA IPC or one process): ONE kernel stages, signals, waits and pulls — no rendezvous, no sync; what
        // follows on the stream sees the gathered tables, and the shards may be dropped right away (stream order)
        PolySet sh{}, os{};
        for (uint32_t k = 0; k < count; k++) {
            sh.src[k] = (const uint32_t *)src[k]->ptr;
            os.dst[k] = dst[k];
        }
        {
            ProfScope _ps(ctx, "gather_xchg", (uint64_t)count * n * ctx->world * sizeof(uint32_t));
            launch_gather_xchg(ctx->d_xchg_view, ctx->world, sh, os, (int)count, n, ++ctx->gather_seq, ctx->d_ticket, ctx->d_mail, ctx->stream);
        }
        rc = check_launch(ctx, "gather_xchg");
        if (rc) drop();
        return rc;
    }
    if (ctx->local) { // same process: every rank reads the peers' shards directly over NVLink; one rendezvous for all tables
        for (uint32_t k = 0; k < count; k++) ctx->local->ptrs[k][ctx->rank] = src[k]->ptr;
        ctx->local->barrier();
        for (uint32_t k = 0; k < count && rc == ZB_OK; k++) {
            PeerSrc ps{};
            for (int q = 0; q < ctx->world; q++) ps.p[q] = (const uint32_t *)ctx->local->ptrs[k][q];
            launch_interleave_peers(ps, dst[k], n, (uint32_t)ctx->world, ctx->sm_count, ctx->stream);
            rc = check_launch(ctx, "interleave_peers");
        }
        if (rc == ZB_OK) rc = zb_sync(ctx);
        else cudaStreamSynchronize(ctx->stream);
        ctx->local->barrier(); // the peers have finished reading this rank's shards
        if (rc) drop();
        return rc;
    }
    if (rc) {
        drop();
        return rc;
    }
    if (!ctx->nccl_comm) {
        drop();
        return ZB_ERR_BAD_ARGUMENT;
    }
    BufRef tmp;
    rc = dev_alloc(ctx, (size_t)count * n * ctx->world * sizeof(uint32_t), &tmp);
    if (rc) {
        drop();
        return rc;
    }
    for (uint32_t k = 0; k < count; k++) {
        uint32_t *t = (uint32_t *)tmp->ptr + (size_t)k * n * ctx->world;
        const int r = g_nccl.AllGather(src[k]->ptr, t, (size_t)n, NCCL_UINT32, ctx->nccl_comm, ctx->stream);
        if (r) {
            cudaStreamSynchronize(ctx->stream);
            drop();
            return nccl_fail(ctx, r, "ncclAllGather");
        }
        launch_interleave(t, dst[k], n, (uint32_t)ctx->world, ctx->stream);
        rc = check_launch(ctx, "interleave");
        if (rc) break;
    }
    if (rc == ZB_OK) rc = zb_sync(ctx);
    else cudaStreamSynchronize(ctx->stream);
    if (rc) drop();
    return rc;
}

int32_t zb_comm_allgather_cyclic(zb_ctx *ctx, zb_mle local, zb_mle *out) {
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    return zb_comm_allgather_cyclic_batch(ctx, &local, 1, out);
}

int32_t zb_comm_p2p_handle(zb_ctx *ctx, uint8_t out[64]) {
    tail_quiesce(ctx);
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
    if (!ctx->d_xchg) {
        const size_t bytes = XCHG_BUFFER_BYTES; // payload sets + the staging area of the in-kernel all-gather
        CK(cudaMalloc(&ctx->d_xchg, bytes));
        CK(cudaMemset(ctx->d_xchg, 0, bytes));
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->d_xchg));
    memcpy(out, &h, 64);
    return ZB_OK;
}

int32_t zb_comm_p2p_attach(zb_ctx *ctx, const uint8_t *handles) {
    tail_quiesce(ctx);
    if (!handles || !ctx->d_xchg || ctx->world < 2 || ctx->world > XCHG_MAX_RANKS || ctx->d_xchg_view) return ZB_ERR_BAD_ARGUMENT;
    XchgView view{};
    view.rank = ctx->rank;
    view.world = ctx->world;
    {
        // how long the last CTA waits for the other ranks' rows. Ranks are only loosely coupled: each one uploads its own
        // shard before it proves, and those uploads share the host's cores and memory, so seconds of skew are normal.
        const char *e = getenv("ZB_XCHG_PATIENCE_S");
        double sec = e && *e ? atof(e) : 30.0;
        if (sec < 0.1) sec = 0.1;
        if (sec > 100.0) sec = 100.0; // stays below the host's 120 s wait for the kernel
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        view.patience = (long long)(sec * 1e3 * (khz > 0 ? khz : 1965000));
    }
    for (int q = 0; q < ctx->world; q++) {
        if (q == ctx->rank) {
            view.peer[q] = ctx->d_xchg;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)q, 64);
        CK(cudaIpcOpenMemHandle(&ctx->xchg_peer[q], h, cudaIpcMemLazyEnablePeerAccess));
        view.peer[q] = (unsigned long long *)ctx->xchg_peer[q];
    }
    CK(cudaMalloc(&ctx->d_xchg_stats, 2 * sizeof(unsigned long long)));
    CK(cudaMemset(ctx->d_xchg_stats, 0, 2 * sizeof(unsigned long long)));
    view.stats = ctx->d_xchg_stats;
    CK(cudaMalloc(&ctx->d_xchg_view, sizeof(XchgView)));
    CK(cudaMemcpy(ctx->d_xchg_view, &view, sizeof(XchgView), cudaMemcpyHostToDevice));
    return ZB_OK;
}

int32_t zb_comm_unique_id(const char *nccl_path, uint8_t out[128]) {
    int32_t rc = nccl_load(nullptr, nccl_path);
    if (rc) return rc;
    NcclId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r) return ZB_ERR_NCCL;
    memcpy(out, id.internal, 128);
    return ZB_OK;
}

int32_t zb_comm_init(zb_ctx *ctx, const char *nccl_path, const uint8_t unique_id[128], int32_t rank, int32_t world) {
    tail_quiesce(ctx);
    if (world < 1 || rank < 0 || rank >= world || (world & (world - 1))) return ZB_ERR_BAD_ARGUMENT;
    if (ctx->nccl_comm) return ZB_ERR_BAD_ARGUMENT;
    int32_t rc = nccl_load(ctx, nccl_path);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    NcclId id;
    memcpy(id.internal, unique_id, 128);
    int r = g_nccl.CommInitRank(&ctx->nccl_comm, world, id, rank);
    if (r) return nccl_fail(ctx, r, "ncclCommInitRank");
    ctx->rank = rank;
    ctx->world = world;
    CK(cudaMalloc(&ctx->d_comm, 64 * sizeof(unsigned long long)));
    CK(cudaHostAlloc((void **)&ctx->h_comm, 64 * sizeof(unsigned long long), cudaHostAllocDefault));
    // first collective pays the connection setup: do it now
    uint64_t warm[1] = {1};
    rc = zb_comm_allreduce_u64(ctx, warm, 1);
    if (rc) return rc;
    return warm[0] == (uint64_t)world ? ZB_OK : ZB_ERR_NCCL;
}

int32_t zb_comm_info(zb_ctx *ctx, int32_t *rank, int32_t *world) {
    // the front of a multi-device context acts as an ordinary one-GPU context for plain handles; it is rank 0 of the
    // group only inside a per-rank job
    const bool plain = group_front(ctx);
    if (rank) *rank = plain ? 0 : ctx->rank;
    if (world) *world = plain ? 1 : ctx->world;
    return ZB_OK;
}

int32_t zb_comm_allreduce_u64(zb_ctx *ctx, uint64_t *vals, uint32_t n) {
    tail_quiesce(ctx);
    if (!vals || n == 0 || n > 64) return ZB_ERR_BAD_ARGUMENT;
    if (ctx->world == 1) return ZB_OK;
    if (!ctx->nccl_comm) return ZB_ERR_BAD_ARGUMENT;
    memcpy(ctx->h_comm, vals, n * sizeof(uint64_t));
    CK(cudaMemcpyAsync(ctx->d_comm, ctx->h_comm, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    int r = g_nccl.AllReduce(ctx->d_comm, ctx->d_comm, n, NCCL_UINT64, NCCL_SUM, ctx->nccl_comm, ctx->stream);
    if (r) return nccl_fail(ctx, r, "ncclAllReduce");
    CK(cudaMemcpyAsync(ctx->h_comm, ctx->d_comm, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(vals, ctx->h_comm, n * sizeof(uint64_t));
    return ZB_OK;
}

int32_t zb_comm_destroy(zb_ctx *ctx) {
    tail_quiesce(ctx);
    if (ctx->d_xchg) {
        cudaStreamSynchronize(ctx->stream);
        ctx->comm_reduce = 0;
        for (int q = 0; q < XCHG_MAX_RANKS; q++)
            if (ctx->xchg_peer[q]) {
                cudaIpcCloseMemHandle(ctx->xchg_peer[q]);
                ctx->xchg_peer[q] = nullptr;
            }
        if (ctx->d_xchg_view) cudaFree(ctx->d_xchg_view);
        ctx->d_xchg_view = nullptr;
        if (ctx->d_xchg_stats) cudaFree(ctx->d_xchg_stats);
        ctx->d_xchg_stats = nullptr;
        cudaFree(ctx->d_xchg);
        ctx->d_xchg = nullptr;
    }
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
        cudaFree(ctx->d_comm);
        cudaFreeHost(ctx->h_comm);
        ctx->d_comm = nullptr;
        ctx->h_comm = nullptr;
        ctx->rank = 0;
        ctx->world = 1;
    }
    return ZB_OK;
}

/* ------------------------------------------------------------------ single-process multi-GPU (zb_ctx_create_mask) */

} // extern "C"

static void group_worker(Group *g, int rank) {
    cudaSetDevice(g->child[rank]->device);
    t_in_group_job = true;
    uint64_t seen = 0;
    for (;;) {
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_go.wait(lk, [&] { return g->stop || g->epoch != seen; });
        if (g->stop) return;
        seen = g->epoch;
        lk.unlock();
        const int32_t rc = g->job(rank);
        lk.lock();
        g->status[rank] = rc;
        if (--g->pending == 0) g->cv_done.notify_one();
    }
}

static int32_t group_run(zb_ctx *front, const std::function<int32_t(int)> &job) {
    Group *g = front->group;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job = job;
        g->pending = g->world - 1;
        g->epoch++;
    }
    g->cv_go.notify_all();
    t_in_group_job = true;
    const int32_t rc0 = job(0);
    t_in_group_job = false;
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    ensure_device(front);
    if (rc0) return rc0;
    for (int r = 1; r < g->world; r++)
        if (g->status[r]) {
            front->last_error = "rank " + std::to_string(r) + ": " + g->child[r]->last_error;
            return g->status[r];
        }
    return ZB_OK;
}

static GMle *get_gmle(zb_ctx *front, zb_mle h) {
    auto it = front->group->mles.find(h);
    return it == front->group->mles.end() ? nullptr : &it->second;
}
static GTree *get_gtree(zb_ctx *front, zb_tree h) {
    auto it = front->group->trees.find(h);
    return it == front->group->trees.end() ? nullptr : &it->second;
}
static int32_t group_unsupported(zb_ctx *front, const char *what) {
    front->last_error = std::string(what) + ": not available on a multi-device context (use one context per device for it)";
    return ZB_ERR_BAD_ARGUMENT;
}
static uint32_t log2_world(int world) {
    uint32_t k = 0;
    while ((1 << k) < world) k++;
    return k;
}
// registers a sharded table made of per-rank tables h[r]
static zb_mle group_new_mle(zb_ctx *front, const zb_mle *h, uint64_t n_total) {
    GMle gm{};
    for (int r = 0; r < front->group->world; r++) gm.child[r] = h[r];
    gm.n = n_total;
    const uint64_t id = GROUP_HANDLE_BIT | front->group->next_handle++;
    front->group->mles[id] = gm;
    return id;
}
static void group_free_children(zb_ctx *front, const zb_mle *h) {
    Group *g = front->group;
    group_run(front, [&](int r) { return h[r] ? zb_mle_free(g->child[r], h[r]) : ZB_OK; });
}

// Every rank creates its cyclic shard with `make(rank, child ctx, &handle)`
static int32_t group_make_mle(zb_ctx *front, uint64_t n_total, zb_mle *out, const std::function<int32_t(int, zb_ctx *, zb_mle *)> &make) {
    Group *g = front->group;
    zb_mle h[XCHG_MAX_RANKS] = {0};
    const int32_t rc = group_run(front, [&](int r) { return make(r, g->child[r], &h[r]); });
    if (rc) {
        group_free_children(front, h);
        return rc;
    }
    *out = group_new_mle(front, h, n_total);
    return ZB_OK;
}

// Host order -> cyclic shards: rank b uploads the CONTIGUOUS slice [b B, (b+1) B) of the host table over its own PCIe link
// (every link carries 1/world of the table exactly once, narrowed to 4 bytes on the way), then deals it out over NVLink:
// element j P + q of the slice is local element b B/P + j of rank q's shard (push, 128-byte coalesced peer stores).
template <typename T>
static int32_t zg_upload(zb_ctx *front, const T *evals, uint64_t n, zb_mle *out) {
    Group *g = front->group;
    const uint64_t P = (uint64_t)g->world;
    int32_t rc = check_pow2(n);
    if (rc) return rc;
    if (!evals || !out) return ZB_ERR_BAD_ARGUMENT;
    if (n < P * P) return group_unsupported(front, "a table with fewer than world^2 entries");
    const uint64_t B = n / P;
    zb_mle shard[XCHG_MAX_RANKS] = {0};
    rc = group_run(front, [&](int r) -> int32_t {
        zb_ctx *ctx = g->child[r];
        tail_quiesce(ctx);
        Mle *m = nullptr;
        int32_t e = new_mle(ctx, B, &shard[r], &m); // (no early return below: every rank must reach both barriers)
        BufRef slice;
        if (e == ZB_OK) e = dev_alloc(ctx, B * sizeof(uint32_t), &slice);
        if (e == ZB_OK) {
            if constexpr (sizeof(T) == 8) e = upload_narrow(ctx, (const uint64_t *)evals + (uint64_t)r * B, B, (uint32_t *)slice->ptr);
            else {
                e = staged_h2d(ctx, slice->ptr, evals + (uint64_t)r * B, B * sizeof(uint32_t));
                if (e == ZB_OK) {
                    launch_check_u32((const uint32_t *)slice->ptr, B, ctx->d_err, ctx->stream);
                    e = check_launch(ctx, "check");
                }
                if (e == ZB_OK) e = read_err_flag(ctx);
            }
        }
        ctx->local->ptr[r] = m ? m->d() : nullptr;
        ctx->local->barrier(); // every rank's shard exists and its address is known (also on a failure: keep the ranks in step)
        bool all = true;
        for (uint64_t q = 0; q < P; q++) all = all && ctx->local->ptr[q] != nullptr;
        if (e == ZB_OK && !all) e = ZB_ERR_OOM; // a peer could not allocate its shard: nobody deals
        if (e == ZB_OK) {
            PeerDst pd{};
            for (uint64_t q = 0; q < P; q++) pd.p[q] = (uint32_t *)ctx->local->ptr[q] + (uint64_t)r * (B / P);
            launch_deal_peers((const uint32_t *)slice->ptr, pd, B / P, (uint32_t)P, ctx->sm_count, ctx->stream);
            e = check_launch(ctx, "deal_peers");
            if (e == ZB_OK) e = zb_sync(ctx);
        }
        ctx->local->barrier(); // all pushes have landed before anybody uses (or frees) a shard
        return e;
    });
    if (rc) {
        group_free_children(front, shard);
        return rc;
    }
    *out = group_new_mle(front, shard, n);
    return ZB_OK;
}

static int32_t zg_mle_free(zb_ctx *front, zb_mle h) {
    GMle *gm = get_gmle(front, h);
    if (!gm) return ZB_ERR_BAD_HANDLE;
    zb_mle c[XCHG_MAX_RANKS];
    memcpy(c, gm->child, sizeof(c));
    front->group->mles.erase(h);
    group_free_children(front, c);
    return ZB_OK;
}

static int32_t zg_mle_download(zb_ctx *front, zb_mle h, uint64_t offset, uint64_t *out, uint64_t n) {
    GMle *gm = get_gmle(front, h);
    if (!gm) return ZB_ERR_BAD_HANDLE;
    if (offset > gm->n || n > gm->n - offset || !out) return ZB_ERR_BAD_ARGUMENT;
    Group *g = front->group;
    const uint64_t P = (uint64_t)g->world, nl = gm->n / P;
    std::vector<std::vector<uint64_t>> tmp(P);
    int32_t rc = group_run(front, [&](int r) {
        tmp[r].resize(nl);
        return zb_mle_download(g->child[r], gm->child[r], tmp[r].data(), nl);
    });
    if (rc) return rc;
    for (uint64_t i = 0; i < n; i++) out[i] = tmp[(offset + i) % P][(offset + i) / P];
    return ZB_OK;
}

// Multilinear.eval on cyclic shards (LSB-first: the rank is index bits 0 .. log2 P - 1, bound by point[0 .. log2 P)):
// value = sum_r eq(point[0..k), r) * eval_r(point[k..v))
static int32_t zg_mle_eval(zb_ctx *front, zb_mle h, const uint64_t *point, uint32_t npoint, uint64_t *out) {
    GMle *gm = get_gmle(front, h);
    if (!gm) return ZB_ERR_BAD_HANDLE;
    Group *g = front->group;
    const uint32_t v = (uint32_t)__builtin_ctzll(gm->n), k = log2_world(g->world);
    if (npoint != v) return ZB_ERR_WRONG_NUM_VARS;
    for (uint32_t i = 0; i < v; i++)
        if (point[i] >= bb::P) return ZB_ERR_NOT_CANONICAL;
    uint64_t part[XCHG_MAX_RANKS] = {0};
    int32_t rc = group_run(front, [&](int r) { return zb_mle_eval(g->child[r], gm->child[r], point + k, v - k, &part[r]); });
    if (rc) return rc;
    uint32_t acc = 0;
    for (int r = 0; r < g->world; r++) {
        uint32_t w = 1;
        for (uint32_t b = 0; b < k; b++) w = bb::mul(w, ((r >> b) & 1) ? (uint32_t)point[b] : bb::sub(1u, (uint32_t)point[b]));
        acc = bb::add(acc, bb::mul(w, (uint32_t)part[r]));
    }
    *out = acc;
    return ZB_OK;
}

static int32_t zg_mle_sum(zb_ctx *front, zb_mle h, uint64_t *out) {
    GMle *gm = get_gmle(front, h);
    if (!gm || !out) return gm ? ZB_ERR_BAD_ARGUMENT : ZB_ERR_BAD_HANDLE;
    Group *g = front->group;
    uint64_t part[XCHG_MAX_RANKS] = {0};
    int32_t rc = group_run(front, [&](int r) { return zb_mle_sum(g->child[r], gm->child[r], &part[r]); });
    if (rc) return rc;
    uint64_t s = 0;
    for (int r = 0; r < g->world; r++) s += part[r];
    *out = s % bb::P;
    return ZB_OK;
}

// SimpleMerkleTree.build for sharded tables: rank r gathers the CONTIGUOUS block [r B, (r+1) B) out of the cyclic shards
// over NVLink (the tree's own copy of the values, merkle_tree.zig:291), builds that subtree, and the host hashes the
// log2(world) levels above the world subtree roots (mergeHashesSHA3, hash.zig:187-195).
static int32_t zg_merkle_build(zb_ctx *front, const zb_mle *polys, uint32_t count, zb_tree *trees, uint8_t *roots) {
    Group *g = front->group;
    const uint64_t P = (uint64_t)g->world;
    if (!polys || !trees || count == 0) return ZB_ERR_BAD_ARGUMENT;
    std::vector<GMle *> gms(count);
    for (uint32_t i = 0; i < count; i++) {
        gms[i] = is_group_handle(polys[i]) ? get_gmle(front, polys[i]) : nullptr;
        if (!gms[i]) return ZB_ERR_BAD_HANDLE;
        if (gms[i]->n != gms[0]->n) return ZB_ERR_DIFFERENT_NUM_VARS;
    }
    const uint64_t n = gms[0]->n, B = n / P;
    if (B < P) return group_unsupported(front, "a table with fewer than world^2 entries");
    std::vector<std::vector<zb_tree>> ct(P, std::vector<zb_tree>(count, 0));
    std::vector<std::vector<uint8_t>> croots(P, std::vector<uint8_t>(32 * (size_t)count));
    int32_t rc = group_run(front, [&](int r) -> int32_t {
        zb_ctx *ctx = g->child[r];
        int32_t e = ZB_OK;
        std::vector<zb_mle> blocks(count, 0);
        for (uint32_t i = 0; i < count; i++) {
            tail_quiesce(ctx);
            Mle *bm = nullptr;
            if (e == ZB_OK) e = new_mle(ctx, B, &blocks[i], &bm);
            Mle *mine = get_mle(ctx, gms[i]->child[r]);
            ctx->local->ptr[r] = mine ? mine->d() : nullptr;
            ctx->local->barrier();
            if (e == ZB_OK && !mine) e = ZB_ERR_BAD_HANDLE;
            if (e == ZB_OK) {
                PeerSrc ps{};
                for (uint64_t q = 0; q < P; q++) ps.p[q] = (const uint32_t *)ctx->local->ptr[q] + (uint64_t)r * (B / P);
                launch_interleave_peers(ps, bm->d(), B / P, (uint32_t)P, ctx->sm_count, ctx->stream);
                e = check_launch(ctx, "interleave_peers");
                if (e == ZB_OK) e = zb_sync(ctx);
            }
            ctx->local->barrier(); // nobody's shard pointer slot is overwritten while a peer still reads it
        }
        if (e == ZB_OK) e = zb_merkle_build(ctx, blocks.data(), count, ct[r].data(), croots[r].data());
        for (uint32_t i = 0; i < count; i++)
            if (blocks[i]) zb_mle_free(ctx, blocks[i]); // the trees hold their own copies
        return e;
    });
    if (rc) {
        group_run(front, [&](int r) {
            for (uint32_t i = 0; i < count; i++)
                if (ct[r][i]) zb_merkle_free(g->child[r], ct[r][i]);
            return ZB_OK;
        });
        return rc;
    }
    const uint32_t k = log2_world(g->world);
    for (uint32_t i = 0; i < count; i++) {
        GTree gt{};
        gt.n_values = n;
        gt.height = (uint32_t)__builtin_ctzll(n);
        gt.top.assign((2 * P - 1) * 32, 0);
        for (uint64_t r = 0; r < P; r++) {
            gt.child[r] = ct[r][i];
            memcpy(&gt.top[32 * r], &croots[r][32 * (size_t)i], 32);
        }
        for (uint32_t l = 0; l < k; l++) {
            const uint64_t in = 2 * P - (2 * P >> l), outo = 2 * P - (2 * P >> (l + 1));
            for (uint64_t j = 0; j < (P >> (l + 1)); j++) zigz::Sha3_256::hash(&gt.top[32 * (in + 2 * j)], 64, &gt.top[32 * (outo + j)]);
        }
        memcpy(gt.root, &gt.top[32 * (2 * P - 2)], 32);
        if (roots) memcpy(roots + 32 * (size_t)i, gt.root, 32);
        const uint64_t id = GROUP_HANDLE_BIT | g->next_handle++;
        g->trees[id] = gt;
        trees[i] = id;
    }
    return ZB_OK;
}

static int32_t zg_merkle_open(zb_ctx *front, zb_tree h, uint64_t index, uint8_t *siblings, uint8_t *dirs, uint64_t *leaf_value) {
    GTree *gt = get_gtree(front, h);
    if (!gt) return ZB_ERR_BAD_HANDLE;
    if (index >= gt->n_values) return ZB_ERR_INDEX_OUT_OF_BOUNDS; // merkle_tree.zig:325
    Group *g = front->group;
    const uint64_t P = (uint64_t)g->world, B = gt->n_values / P;
    const uint32_t k = log2_world(g->world), hl = gt->height - k;
    const uint64_t owner = index / B;
    // the owner's context serves the path inside its subtree (a plain call from this thread: contexts are thread-agnostic)
    int32_t rc = zb_merkle_open(g->child[owner], gt->child[owner], index % B, siblings, dirs, leaf_value);
    ensure_device(front);
    if (rc) return rc;
    for (uint32_t l = 0; l < k; l++) { // :341-345 continued above the subtree roots
        const uint64_t pos = owner >> l, off = 2 * P - (2 * P >> l);
        memcpy(siblings + 32 * (size_t)(hl + l), &gt->top[32 * (off + (pos ^ 1))], 32);
        dirs[hl + l] = (uint8_t)(pos & 1);
    }
    return ZB_OK;
}

static int32_t zg_merkle_free(zb_ctx *front, zb_tree h) {
    GTree *gt = get_gtree(front, h);
    if (!gt) return ZB_ERR_BAD_HANDLE;
    Group *g = front->group;
    zb_tree c[XCHG_MAX_RANKS];
    memcpy(c, gt->child, sizeof(c));
    g->trees.erase(h);
    group_run(front, [&](int r) { return zb_merkle_free(g->child[r], c[r]); });
    return ZB_OK;
}

extern "C" {

int32_t zb_ctx_create_mask(uint32_t device_mask, zb_ctx **out) {
    if (!out) return ZB_ERR_BAD_ARGUMENT;
    *out = nullptr;
    int devs[XCHG_MAX_RANKS], world = 0;
    for (int d = 0; d < 32; d++)
        if ((device_mask >> d) & 1u) {
            if (world == XCHG_MAX_RANKS) return ZB_ERR_BAD_ARGUMENT;
            devs[world++] = d;
        }
    if (world == 0 || (world & (world - 1))) return ZB_ERR_BAD_ARGUMENT; // 1, 2, 4, 8 or 16 devices
    if (world == 1) return zb_ctx_create(devs[0], out);
    zb_ctx *c[XCHG_MAX_RANKS] = {nullptr};
    auto fail = [&](int32_t rc) {
        for (int r = world - 1; r >= 0; r--)
            if (c[r]) {
                c[r]->group = nullptr;
                zb_ctx_destroy(c[r]);
            }
        return rc;
    };
    for (int r = 0; r < world; r++) {
        const int32_t rc = zb_ctx_create(devs[r], &c[r]);
        if (rc) return fail(rc);
    }
    // peer access between every pair (NVLink / NVSwitch): kernels of one GPU read and write the others' memory directly
    for (int r = 0; r < world; r++) {
        cudaSetDevice(devs[r]);
        for (int q = 0; q < world; q++) {
            if (q == r) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[r], devs[q]);
            if (!can) {
                c[0]->last_error = "no peer access between the selected devices";
                fprintf(stderr, "zigz_b200: device %d cannot access device %d\n", devs[r], devs[q]);
                return fail(ZB_ERR_CUDA);
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ZB_ERR_CUDA);
            cudaGetLastError();
        }
    }
    auto lc = std::make_shared<LocalComm>();
    lc->world = world;
    const size_t xbytes = XCHG_BUFFER_BYTES;
    for (int r = 0; r < world; r++) {
        cudaSetDevice(devs[r]);
        c[r]->local = lc;
        c[r]->rank = r;
        c[r]->world = world;
        if (cudaMalloc(&c[r]->d_xchg, xbytes) != cudaSuccess || cudaMemset(c[r]->d_xchg, 0, xbytes) != cudaSuccess) return fail(ZB_ERR_OOM);
        if (cudaMalloc(&c[r]->d_xchg_stats, 2 * sizeof(unsigned long long)) != cudaSuccess) return fail(ZB_ERR_OOM);
        cudaMemset(c[r]->d_xchg_stats, 0, 2 * sizeof(unsigned long long));
    }
    for (int r = 0; r < world; r++) {
        cudaSetDevice(devs[r]);
        XchgView view{};
        view.rank = r;
        view.world = world;
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, devs[r]);
        view.patience = (long long)(30.0 * 1e3 * (khz > 0 ? khz : 1965000));
        view.stats = c[r]->d_xchg_stats;
        for (int q = 0; q < world; q++) view.peer[q] = c[q]->d_xchg; // plain device pointers: one address space, peer access on
        if (cudaMalloc(&c[r]->d_xchg_view, sizeof(XchgView)) != cudaSuccess) return fail(ZB_ERR_OOM);
        cudaMemcpy(c[r]->d_xchg_view, &view, sizeof(XchgView), cudaMemcpyHostToDevice);
    }
    Group *g = new Group();
    g->world = world;
    for (int r = 0; r < world; r++) g->child[r] = c[r];
    c[0]->group = g;
    for (int r = 1; r < world; r++) g->workers.emplace_back(group_worker, g, r);
    cudaSetDevice(devs[0]);
    *out = c[0];
    return ZB_OK;
}

/* how many device contexts stand behind ctx (1 for an ordinary context) */
int32_t zb_group_size(zb_ctx *ctx) { return ctx && ctx->group ? ctx->group->world : 1; }

/* host twins: run fn(rank's context, rank, world, user) once per device of a multi-device context, concurrently, each on
 * the thread that owns that device; returns the first failing status. On an ordinary context: fn(ctx, 0, 1, user). */
int32_t zb_group_run(zb_ctx *ctx, zb_rank_fn fn, void *user) {
    if (!ctx || !fn) return ZB_ERR_BAD_ARGUMENT;
    if (!group_front(ctx)) return fn(ctx, ctx->rank, ctx->world, user);
    Group *g = ctx->group;
    return group_run(ctx, [&](int r) { return fn(g->child[r], r, g->world, user); });
}

/* rank's shard of a sharded table (a handle of that rank's context) */
int32_t zb_group_mle(zb_ctx *ctx, zb_mle h, int32_t rank, zb_mle *out) {
    if (!ctx || !ctx->group || !out || rank < 0 || rank >= ctx->group->world) return ZB_ERR_BAD_ARGUMENT;
    auto it = ctx->group->mles.find(h);
    if (it == ctx->group->mles.end()) return ZB_ERR_BAD_HANDLE;
    *out = it->second.child[rank];
    return ZB_OK;
}

/* after a consuming prove the shards have been folded away: the sharded table shrinks to `n` entries in total */
int32_t zb_group_mle_set_len(zb_ctx *ctx, zb_mle h, uint64_t n) {
    if (!ctx || !ctx->group) return ZB_ERR_BAD_ARGUMENT;
    auto it = ctx->group->mles.find(h);
    if (it == ctx->group->mles.end()) return ZB_ERR_BAD_HANDLE;
    it->second.n = n;
    return ZB_OK;
}

} // extern "C"
