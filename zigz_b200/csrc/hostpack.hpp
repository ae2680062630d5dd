// Host-side marshalling for the upload path: the reference keeps field elements as 8-byte `struct { value: u64 }`
// (src/core/field.zig:26-27) while the device stores canonical u32. Narrowing on the host with a few threads halves
// the bytes that cross PCIe (the e2e bound of the path) and overlaps with the copies.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace zigz {

// dst[i] = (uint32_t)src[i]; returns true if any src[i] >= p (not canonical)
bool narrow_u64_to_u32(const uint64_t *src, uint32_t *dst, size_t n, uint64_t p);

// minimal persistent fork-join pool: run(fn) calls fn(tid) on every worker and on the caller (tid 0).
// Workers (and the caller waiting for them) spin for a few tens of microseconds before they block on a condition variable:
// the upload path issues one short job per staging chunk (~0.75 ms); an empty fork-join drops from 20-120 us to 1-10 us.
// (On the pool's 16-core hosts the packing itself is memory-bound, so the upload rate did not move: 86.9 GB/s either way.)
class HostPool {
  public:
    explicit HostPool(int threads);
    ~HostPool();
    int size() const { return nthreads_; }
    void run(const std::function<void(int)> &fn);

  private:
    void worker(int tid);
    int nthreads_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    const std::function<void(int)> *job_ = nullptr;
    std::atomic<uint64_t> epoch_{0};
    std::atomic<int> pending_{0};
    std::atomic<bool> stop_{false};
};

} // namespace zigz
