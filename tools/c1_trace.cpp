// Timeline of one d=1 prove through the C ABI (no Python in the loop): per-call host timestamps.
// build: g++ -O2 -std=c++17 -Iinclude tools/c1_trace.cpp -Lzigz_b200 -lzigz_b200 -Wl,-rpath,$PWD/zigz_b200 -o /tmp/c1_trace
#include "zigz_b200.h"
#include "zigz_host.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }
int main(int argc, char **argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 20, reps = argc > 2 ? atoi(argv[2]) : 200;
    zb_ctx *ctx = nullptr;
    if (zb_ctx_create(0, &ctx)) return 1;
    zb_mle p = 0;
    if (zb_mle_synthetic(ctx, 0x5A49475A, 0, 1, 1ull << lg, &p)) return 2;
    std::vector<uint64_t> rp(2 * lg), fp(lg);
    uint64_t fe = 0, cs = 0;
    for (int i = 0; i < 20; i++) zh_sumcheck_prove(ctx, p, rp.data(), fp.data(), &fe, &cs);
    auto t0 = clk::now();
    for (int i = 0; i < reps; i++) zh_sumcheck_prove(ctx, p, rp.data(), fp.data(), &fe, &cs);
    auto t1 = clk::now();
    printf("zh_sumcheck_prove 2^%d: %.2f us per prove (C ABI, %d reps), final_eval %llu\n", lg, us(t0, t1) / reps, reps, (unsigned long long)fe);
    // the pieces
    uint64_t S[1024], r[5] = {1, 2, 3, 4, 5};
    double a = 0, b = 0, c = 0, f = 0;
    for (int i = 0; i < reps; i++) {
        auto s0 = clk::now();
        zb_mle_block_sums(ctx, p, 5, S);
        auto s1 = clk::now();
        zb_mle q = 0, q2 = 0;
        zb_mle_fold_multi(ctx, p, 5, r, &q, 5, S);
        auto s2 = clk::now();
        if (lg - 10 <= 10) zb_mle_fold_multi(ctx, q, 5, r, nullptr, lg - 10, S);
        else zb_mle_fold_multi(ctx, q, 5, r, nullptr, 5, S);
        auto s3 = clk::now();
        zb_mle_free(ctx, q);
        (void)q2;
        auto s4 = clk::now();
        a += us(s0, s1), b += us(s1, s2), c += us(s2, s3), f += us(s3, s4);
    }
    printf("  block_sums %.2f us, fold_multi(out) %.2f us, fold_multi(in place%s) %.2f us, free %.2f us\n", a / reps, b / reps,
           lg - 10 <= 10 ? ", table published" : "", c / reps, f / reps);
    // transcript cost
    zh_transcript *t = zh_transcript_new();
    auto h0 = clk::now();
    uint64_t co[2] = {5, 7}, acc = 0;
    for (int i = 0; i < 1000; i++) {
        zh_transcript_append_fields(t, co, 2);
        acc += zh_transcript_challenge(t);
    }
    auto h1 = clk::now();
    printf("  transcript: %.3f us per round (append 2 + challenge), checksum %llu\n", us(h0, h1) / 1000, (unsigned long long)acc);
    zh_transcript_free(t);
    zb_ctx_destroy(ctx);
    return 0;
}
