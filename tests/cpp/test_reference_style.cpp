// C++ parity test in the shape of the reference's own unit tests (test names follow /root/reference), written against
// include/zigz_host.hpp (the C++ mirror of src/lib.zig) and checked against the C oracle (test infrastructure).
// Built and run by tests/test_gpu_cpp_api.py; compiled (not run) by the CPU suite.
#include "../../include/zigz_host.hpp"
#include "../../oracle/zigz_oracle.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace zigz;
static int failures = 0;
#define EXPECT(cond)                                                         \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            failures++;                                                      \
        }                                                                    \
    } while (0)
template <typename Fn>
static bool throws(const char *name, Fn fn) {
    try {
        fn();
    } catch (const Error &e) {
        return std::strcmp(e.what() + 6, name) == 0; // "error.<Name>"
    }
    return false;
}
static const uint64_t P = ZB_BABYBEAR_P;
static std::vector<F> synth(uint64_t seed, size_t n) {
    std::vector<F> v(n);
    zo_fill_synthetic(P, seed, 0, n, v.data());
    return v;
}

int main() {
    Context ctx(0);
    { // "multilinear: evaluation at boolean points" / "partial evaluation" / "round polynomial" (multilinear.zig:383-506)
        auto p = Multilinear::init(ctx, {1, 2, 3, 4});
        EXPECT(p.num_vars() == 2);
        EXPECT(p.eval({0, 0}) == 1 && p.eval({1, 0}) == 2 && p.eval({0, 1}) == 3 && p.eval({1, 1}) == 4);
        auto q = p.partialEval(0);
        EXPECT(q.num_vars() == 1 && q.evaluations() == std::vector<F>({1, 2}));
        EXPECT(p.sumOverHypercube() == 10);
        auto c = p.roundPolynomial();
        EXPECT(c[0] == 3 && c[1] == 4);
        EXPECT(throws("LengthNotPowerOfTwo", [&] { Multilinear::init(ctx, {1, 2, 3}); }));
        EXPECT(throws("WrongNumberOfVariables", [&] { p.eval({1}); }));
    }
    { // SumcheckProver.prove against the oracle, sizes 2^1 .. 2^16; toBytes identical
        for (int lg : {1, 2, 5, 11, 16}) {
            auto e = synth(0x5A49475A, (size_t)1 << lg);
            auto poly = Multilinear::init(ctx, e);
            SumcheckProof pr = SumcheckProver::prove(poly);
            std::vector<uint64_t> rp(2 * lg), pt(lg);
            uint64_t fe, cs;
            EXPECT(zo_sumcheck_prove(P, e.data(), e.size(), rp.data(), pt.data(), &fe, &cs) == 0);
            std::vector<uint8_t> want((2 + 3 * lg) * 8);
            zo_sumcheck_proof_to_bytes(lg, rp.data(), pt.data(), fe, want.data());
            EXPECT(pr.toBytes() == want);
            EXPECT(poly.evaluations() == e); // prove leaves the polynomial untouched
        }
        auto one = Multilinear::init(ctx, {5});
        EXPECT(throws("NoVariables", [&] { SumcheckProver::prove(one); }));
        auto p8 = Multilinear::init(ctx, synth(1, 8));
        EXPECT(throws("WrongNumberOfChallenges", [&] { SumcheckProver::proveInteractive(p8, {1, 2}); }));
    }
    { // "merkle_tree: valid proof verifies" / "invalid proof rejected" / "non-power-of-2 padding" (merkle_tree.zig:470-529)
        auto t = SimpleMerkleTree::build(ctx, {1, 2, 3, 4, 5});
        EXPECT(t.height() == 3);
        uint8_t root[32];
        uint64_t vals[5] = {1, 2, 3, 4, 5};
        EXPECT(zo_merkle_build(vals, 5, nullptr, root, nullptr) == 0);
        EXPECT(std::memcmp(root, t.getRoot().data(), 32) == 0);
        for (size_t i = 0; i < 5; i++) {
            auto pr = t.open(i);
            EXPECT(pr.value == vals[i] && pr.path.siblings.size() == 3);
            EXPECT(SimpleMerkleTree::verify(t.getRoot(), pr));
            pr.value = (pr.value + 1) % P;
            EXPECT(!SimpleMerkleTree::verify(t.getRoot(), pr));
        }
        EXPECT(throws("IndexOutOfBounds", [&] { t.open(5); }));
        EXPECT(throws("EmptyValues", [&] { SimpleMerkleTree::build(ctx, {}); }));
    }
    { // "polynomial_commit: open and verify" (polynomial_commit.zig:300-320)
        auto e = synth(77, 1 << 10);
        auto poly = Multilinear::init(ctx, e);
        auto ct = CommitmentScheme::commit(poly);
        auto pt = synth(78, 10);
        auto op = CommitmentScheme::open(poly, ct.second, pt);
        uint64_t want;
        EXPECT(zo_mle_eval(P, e.data(), e.size(), pt.data(), 10, &want) == 0 && op.value == want);
        EXPECT(op.merkle_proof.index == pt[0] % 1024 && op.merkle_proof.value == e[op.merkle_proof.index]);
        EXPECT(CommitmentScheme::verify(ct.first, op));
        EXPECT(throws("PointDimensionMismatch", [&] { CommitmentScheme::open(poly, ct.second, {1, 2}); }));
    }
    { // "lasso_prover: proof with mapping" / "invalid mapping detection" (lasso_prover.zig:352-412), XOR 2-bit table
        std::vector<F> table(16 * 3);
        zo_build_table(P, ZO_TABLE_XOR, 2, table.data());
        std::vector<F> q = {3, 2, 1, 0, 0, 0};
        LassoProof pr = LassoProver::proveWithMapping(ctx, table, q, {14, 0}, 3);
        EXPECT(pr.num_lookups == 2 && pr.sumcheck_proof.num_vars == 1);
        uint64_t rp[2], pt[1], fe;
        uint32_t nv;
        uint8_t qc[32], tc[32];
        EXPECT(zo_lasso_prove(P, table.data(), 16, q.data(), 2, 3, rp, pt, &fe, &nv, qc, tc) == 0);
        EXPECT(pr.sumcheck_proof.round_polynomials[0][0] == rp[0] && pr.sumcheck_proof.final_eval == fe);
        EXPECT(std::memcmp(qc, pr.query_commitment.data(), 32) == 0 && std::memcmp(tc, pr.table_commitment.data(), 32) == 0);
        EXPECT(throws("QueryTableMismatch", [&] { LassoProver::proveWithMapping(ctx, table, q, {13, 0}, 3); }));
        EXPECT(throws("InvalidMapping", [&] { LassoProver::proveWithMapping(ctx, table, q, {16, 0}, 3); }));
        EXPECT(throws("NoQueries", [&] { LassoProver::prove(ctx, table, {}, 3); }));
    }
    std::printf(failures ? "cpp api: %d FAILURES\n" : "cpp api: all checks passed\n", failures);
    return failures ? 1 : 0;
}
