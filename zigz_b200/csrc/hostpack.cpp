#include "hostpack.hpp"

namespace zigz {

bool narrow_u64_to_u32(const uint64_t *src, uint32_t *dst, size_t n, uint64_t p) {
    uint64_t bad = 0;
    // plain loop: gcc -O3 -march=x86-64-v3 turns it into 256-bit loads + vpermd/vpshufd packs
    for (size_t i = 0; i < n; i++) {
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
        stop_ = true;
    }
    cv_start_.notify_all();
    for (auto &t : threads_) t.join();
}

void HostPool::worker(int tid) {
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(int)> *job;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_start_.wait(lk, [&] { return stop_ || epoch_ != seen; });
            if (stop_) return;
            seen = epoch_;
            job = job_;
        }
        (*job)(tid);
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) cv_done_.notify_one();
        }
    }
}

void HostPool::run(const std::function<void(int)> &fn) {
    if (nthreads_ == 1) {
        fn(0);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_ = &fn;
        pending_ = nthreads_ - 1;
        epoch_++;
    }
    cv_start_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
}

} // namespace zigz
