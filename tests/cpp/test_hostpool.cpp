// Fork-join pool and u64 -> u32 narrowing of the upload path (zigz_b200/csrc/hostpack.*): host-only, no GPU needed.
// Every worker must run every job exactly once, across spin and sleep phases, and the pool must start and stop cleanly.
#include "../../zigz_b200/csrc/hostpack.hpp"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

int main() {
    for (int T : {1, 2, 3, 8}) {
        zigz::HostPool pool(T);
        std::vector<std::atomic<long>> cnt(T);
        for (auto &c : cnt) c = 0;
        const long runs = 20000;
        for (long r = 0; r < runs; r++) {
            pool.run([&](int tid) { cnt[tid].fetch_add(1, std::memory_order_relaxed); });
            if (r % 4001 == 0) std::this_thread::sleep_for(std::chrono::microseconds(400)); // let the workers fall asleep
        }
        for (auto &c : cnt)
            if (c.load() != runs) {
                printf("T=%d: a worker ran %ld of %ld jobs\n", T, c.load(), runs);
                return 1;
            }
    }
    for (int k = 0; k < 50; k++) { // construction / destruction, with and without work, awake and asleep
        zigz::HostPool p(4);
        if (k & 1) p.run([](int) {});
        if (k % 10 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    // narrowing: values, canonical check (>= p and top-bit cases), ragged lengths and unaligned destinations
    const uint64_t P = 2013265921ull;
    std::vector<uint64_t> src(1000);
    for (size_t i = 0; i < src.size(); i++) src[i] = (i * 2654435761ull) % P;
    alignas(64) static uint32_t dst[1008];
    for (size_t n : std::vector<size_t>{0, 1, 7, 8, 9, 64, 999, 1000})
        for (size_t off : std::vector<size_t>{0, 1, 8}) {
            if (zigz::narrow_u64_to_u32(src.data(), dst + off, n, P)) return 2;
            for (size_t i = 0; i < n; i++)
                if (dst[off + i] != (uint32_t)src[i]) return 3;
        }
    for (uint64_t bad : std::vector<uint64_t>{P, P + 1, 1ull << 31, 1ull << 32, ~0ull, 1ull << 63})
        for (size_t pos : std::vector<size_t>{0, 5, 8, 999}) {
            std::vector<uint64_t> s2 = src;
            s2[pos] = bad;
            if (!zigz::narrow_u64_to_u32(s2.data(), dst, s2.size(), P)) {
                printf("non-canonical %llu at %zu not flagged\n", (unsigned long long)bad, pos);
                return 4;
            }
        }
    printf("hostpool ok\n");
    return 0;
}
