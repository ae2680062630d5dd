#include "hostpack.hpp"

#include <cstdlib>
#include <immintrin.h>

namespace zigz {

bool narrow_u64_to_u32(const uint64_t *src, uint32_t *dst, size_t n, uint64_t p) {
    uint64_t bad = 0;
    size_t i = 0;
#if defined(__AVX2__)
    // 8 elements per step: two 256-bit loads, keep the low dwords, one 256-bit NON-TEMPORAL store (the staging buffer
    // is written once and read by the DMA engine: bypassing the cache saves the read-for-ownership traffic)
    static const bool use_nt = [] {
        const char *e = getenv("ZB_PACK_NT");
        return !(e && *e == '0');
    }();
    if (use_nt && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
        const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
        const __m256i pm1 = _mm256_set1_epi64x((long long)(p - 1));
        __m256i over = _mm256_setzero_si256();
        for (; i + 8 <= n; i += 8) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 4));
            // canonical values are < 2^31, so a signed 64-bit compare against p - 1 is exact for the check (anything with
            // the top bit set compares "less", which the high-dword test below catches)
            over = _mm256_or_si256(over, _mm256_or_si256(_mm256_cmpgt_epi64(a, pm1), _mm256_cmpgt_epi64(b, pm1)));
            over = _mm256_or_si256(over, _mm256_or_si256(_mm256_srli_epi64(a, 63), _mm256_srli_epi64(b, 63)));
            const __m256i lo = _mm256_permutevar8x32_epi32(a, pick); // a0 a1 a2 a3 | a0 a1 a2 a3
            const __m256i hi = _mm256_permutevar8x32_epi32(b, pick);
            _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_blend_epi32(lo, hi, 0xF0));
        }
        _mm_sfence();
        bad |= (uint64_t)!_mm256_testz_si256(over, over);
    }
#endif
    for (; i < n; i++) {
        const uint64_t v = src[i];
        bad |= (uint64_t)(v >= p);
        dst[i] = (uint32_t)v;
    }
    return bad != 0;
}

HostPool::HostPool(int threads) : nthreads_(threads < 1 ? 1 : threads) {
    for (int t = 1; t < nthreads_; t++) threads_.emplace_back([this, t] { worker(t); });
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_.store(true, std::memory_order_release);
    }
    cv_start_.notify_all();
    for (auto &t : threads_) t.join();
}

namespace {
constexpr int SPIN_LIMIT = 2000; // x _mm_pause (~40-140 cycles each): tens of microseconds
}

void HostPool::worker(int tid) {
    uint64_t seen = 0;
    for (;;) {
        uint64_t e;
        int spins = 0;
        while ((e = epoch_.load(std::memory_order_acquire)) == seen && !stop_.load(std::memory_order_acquire)) {
            if (++spins < SPIN_LIMIT) {
                _mm_pause();
                continue;
            }
            std::unique_lock<std::mutex> lk(mu_);
            cv_start_.wait(lk, [&] { return stop_.load(std::memory_order_acquire) || epoch_.load(std::memory_order_acquire) != seen; });
        }
        if (e == seen) return; // stop
        seen = e;
        (*job_)(tid); // published before the epoch moved
        if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> lk(mu_); // the caller is either before its predicate check or waiting
            cv_done_.notify_one();
        }
    }
}

void HostPool::run(const std::function<void(int)> &fn) {
    if (nthreads_ == 1) {
        fn(0);
        return;
    }
    job_ = &fn;
    pending_.store(nthreads_ - 1, std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lk(mu_);
        epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_start_.notify_all();
    fn(0);
    int spins = 0;
    while (pending_.load(std::memory_order_acquire) != 0) {
        if (++spins < SPIN_LIMIT) {
            _mm_pause();
            continue;
        }
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
    }
}

} // namespace zigz
