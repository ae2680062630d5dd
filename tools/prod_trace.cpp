// Timeline of one degree-3 product prove through the C ABI (no Python in the loop): where the microseconds of a small prove go.
// build: g++ -O2 -std=c++17 -Iinclude tools/prod_trace.cpp -Lzigz_b200 -lzigz_b200 -Wl,-rpath,$PWD/zigz_b200 -o /tmp/prod_trace
#include "zigz_b200.h"
#include "zigz_host.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }
int main(int argc, char **argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 20, reps = argc > 2 ? atoi(argv[2]) : 200, d = argc > 3 ? atoi(argv[3]) : 3;
    zb_ctx *ctx = nullptr;
    if (zb_ctx_create(0, &ctx)) return 1;
    zb_mle p[3] = {0, 0, 0};
    for (int k = 0; k < d; k++)
        if (zb_mle_synthetic(ctx, 1 + k, 0, 1, 1ull << lg, &p[k])) return 2;
    std::vector<uint64_t> rp(4 * lg), fp(lg);
    uint64_t fe[3] = {0, 0, 0}, cs = 0;
    for (int i = 0; i < 20; i++) zh_prodcheck_prove(ctx, p, d, rp.data(), fp.data(), fe, &cs);
    auto t0 = clk::now();
    for (int i = 0; i < reps; i++) zh_prodcheck_prove(ctx, p, d, rp.data(), fp.data(), fe, &cs);
    auto t1 = clk::now();
    printf("zh_prodcheck_prove d=%d 2^%d: %.2f us per prove (C ABI, %d reps), final_eval %llu\n", d, lg, us(t0, t1) / reps, reps,
           (unsigned long long)fe[0]);
    // the pieces, each call timed on its own
    uint64_t co[4], grid[16], r[2] = {12345, 67890};
    std::vector<double> acc(40, 0.0);
    std::vector<uint32_t> tables(3u << 12);
    int nsteps = 0;
    for (int i = 0; i < reps; i++) {
        int s = 0;
        auto a = clk::now();
        zb_prod_round_coeffs(ctx, p, d, co);
        auto b = clk::now();
        acc[s++] += us(a, b);
        zb_mle q[3] = {0, 0, 0};
        a = clk::now();
        zb_prod_fold_grid(ctx, p, d, 1, r, q, grid);
        b = clk::now();
        acc[s++] += us(a, b);
        uint64_t n = 1ull << (lg - 1);
        while ((n >> 2) > 1024) {
            a = clk::now();
            zb_prod_fold_grid(ctx, q, d, 2, r, nullptr, grid);
            b = clk::now();
            acc[s++] += us(a, b);
            n >>= 2;
        }
        a = clk::now();
        zb_prod_fold_dump(ctx, q, d, 2, r, tables.data());
        b = clk::now();
        acc[s++] += us(a, b);
        n >>= 2;
        zh_transcript *t = zh_transcript_new();
        a = clk::now();
        zh_prodcheck_finish_small(d, tables.data(), n, 0, t, nullptr, rp.data(), fp.data(), fe);
        b = clk::now();
        acc[s++] += us(a, b);
        zh_transcript_free(t);
        a = clk::now();
        for (int k = 0; k < d; k++) zb_mle_free(ctx, q[k]);
        b = clk::now();
        acc[s++] += us(a, b);
        nsteps = s;
    }
    printf("  round_coeffs %.2f | fold1+grid(out) %.2f | fold2+grid:", acc[0] / reps, acc[1] / reps);
    for (int s = 2; s < nsteps - 3; s++) printf(" %.2f", acc[s] / reps);
    printf(" | fold2+dump %.2f | host finish %.2f | free %.2f  (us)\n", acc[nsteps - 3] / reps, acc[nsteps - 2] / reps, acc[nsteps - 1] / reps);
    zb_ctx_destroy(ctx);
    return 0;
}
