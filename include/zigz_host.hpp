// zigz_host.hpp — header-only C++ mirror of the reference's prover-side API (src/lib.zig:9-17) on top of the C ABI.
//
// The reference's host language is Zig (absent from the build image); this is the same surface in C++, one class per
// reference type, same method names, Zig error unions -> `zigz::Error` carrying the reference's error name.
// Everything forwards to zb_* / zh_* (include/zigz_b200.h, include/zigz_host.h); nothing here computes.
//
//   reference (Zig)                                   here
//   Multilinear(F).init / eval / partialEval / ...    zigz::Multilinear
//   FiatShamirTranscript                              zigz::FiatShamirTranscript
//   SumcheckProver(F).prove / proveInteractive        zigz::SumcheckProver
//   SumcheckProof(F) (+ toBytes)                      zigz::SumcheckProof
//   SimpleMerkleTree(F, SHA3Hasher)                   zigz::SimpleMerkleTree
//   CommitmentScheme(F, SHA3Hasher)                   zigz::CommitmentScheme
//   LassoProver(F)                                    zigz::LassoProver
#pragma once
#include "zigz_host.h"

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace zigz {

using F = uint64_t; // canonical BabyBear value, the reference's `struct { value: u64 }` (src/core/field.zig:26-27)
using Digest = std::array<uint8_t, 32>;

struct Error : std::runtime_error { // a Zig error name of the reference, e.g. "NoVariables"
    int32_t code;
    explicit Error(int32_t c) : std::runtime_error(std::string("error.") + zb_status_name(c)), code(c) {}
};
inline void check(int32_t rc) {
    if (rc != ZB_OK) throw Error(rc);
}

class Context { // one GPU + stream; the allocator argument of the reference's calls becomes this handle
  public:
    explicit Context(int device = 0) { check(zb_ctx_create(device, &ctx_)); }
    ~Context() { zb_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    zb_ctx *get() const { return ctx_; }

  private:
    zb_ctx *ctx_ = nullptr;
};

// src/poly/multilinear.zig:20
class Multilinear {
  public:
    static Multilinear init(Context &c, const std::vector<F> &evaluations) { // :36-54
        zb_mle h = 0;
        check(zb_mle_upload(c.get(), evaluations.data(), evaluations.size(), &h));
        return Multilinear(c, h);
    }
    static Multilinear zero(Context &c, uint32_t num_vars) { return constant(c, num_vars, 0); } // :57-70
    static Multilinear constant(Context &c, uint32_t num_vars, F value) {                       // :73-86
        zb_mle h = 0;
        check(zb_mle_constant(c.get(), num_vars, value, &h));
        return Multilinear(c, h);
    }
    Multilinear(Multilinear &&o) noexcept : c_(o.c_), h_(o.h_) { o.h_ = 0; }
    Multilinear(const Multilinear &) = delete;
    ~Multilinear() { deinit(); }
    void deinit() { // :89-91
        if (h_) zb_mle_free(c_->get(), h_);
        h_ = 0;
    }
    uint32_t num_vars() const {
        uint32_t v = 0;
        check(zb_mle_len(c_->get(), h_, nullptr, &v));
        return v;
    }
    std::vector<F> evaluations() const {
        uint64_t n = 0;
        check(zb_mle_len(c_->get(), h_, &n, nullptr));
        std::vector<F> out(n);
        check(zb_mle_download(c_->get(), h_, out.data(), n));
        return out;
    }
    F eval(const std::vector<F> &point) const { // :110-144
        F out = 0;
        check(zb_mle_eval(c_->get(), h_, point.data(), (uint32_t)point.size(), &out));
        return out;
    }
    Multilinear partialEval(F r) const { // :154-180
        zb_mle h = 0;
        check(zb_mle_partial_eval(c_->get(), h_, r, &h, nullptr));
        return Multilinear(*c_, h);
    }
    F sumOverHypercube() const { // :188-194
        F out = 0;
        check(zb_mle_sum(c_->get(), h_, &out));
        return out;
    }
    std::array<F, 2> roundPolynomial() const { // :205-232 -> [s0, s1 - s0]
        uint64_t s[2];
        check(zb_mle_round_sums(c_->get(), h_, s));
        return {s[0], zh_f_sub(s[1], s[0])};
    }
    Multilinear add(const Multilinear &o) const { // :235-250
        zb_mle h = 0;
        check(zb_mle_add(c_->get(), h_, o.h_, &h));
        return Multilinear(*c_, h);
    }
    Multilinear scalarMul(F s) const { // :253-264
        zb_mle h = 0;
        check(zb_mle_scalar_mul(c_->get(), h_, s, &h));
        return Multilinear(*c_, h);
    }
    zb_mle handle() const { return h_; }
    Context &context() const { return *c_; }

  private:
    Multilinear(Context &c, zb_mle h) : c_(&c), h_(h) {}
    Context *c_;
    zb_mle h_;
    friend class LassoProver;
};

// src/core/hash.zig:255-324
class FiatShamirTranscript {
  public:
    FiatShamirTranscript() : t_(zh_transcript_new()) {}
    ~FiatShamirTranscript() { zh_transcript_free(t_); }
    FiatShamirTranscript(const FiatShamirTranscript &) = delete;
    void appendFieldElement(F v) { zh_transcript_append_field(t_, v); }
    void appendFieldElements(const std::vector<F> &v) { zh_transcript_append_fields(t_, v.data(), v.size()); }
    void appendBytes(const void *d, size_t n) { zh_transcript_append_bytes(t_, d, n); }
    F challenge() { return zh_transcript_challenge(t_); }
    zh_transcript *get() const { return t_; }

  private:
    zh_transcript *t_;
};

// src/proofs/sumcheck_protocol.zig:24-109
struct SumcheckProof {
    std::vector<std::array<F, 2>> round_polynomials;
    std::vector<F> final_point;
    F final_eval = 0;
    size_t num_vars = 0;
    std::vector<uint8_t> toBytes() const { // :76-109
        std::vector<uint8_t> out((2 + 3 * num_vars) * 8);
        zh_sumcheck_proof_to_bytes((uint32_t)num_vars, num_vars ? round_polynomials[0].data() : nullptr, final_point.data(),
                                   final_eval, out.data());
        return out;
    }
};

// src/proofs/sumcheck_prover.zig
struct SumcheckProver {
    static SumcheckProof prove(const Multilinear &poly) { // :26-91
        SumcheckProof p;
        p.num_vars = poly.num_vars();
        p.round_polynomials.resize(p.num_vars ? p.num_vars : 1);
        p.final_point.resize(p.num_vars ? p.num_vars : 1);
        check(zh_sumcheck_prove(poly.context().get(), poly.handle(), p.round_polynomials[0].data(), p.final_point.data(),
                                &p.final_eval, nullptr));
        p.round_polynomials.resize(p.num_vars);
        p.final_point.resize(p.num_vars);
        return p;
    }
    static SumcheckProof proveInteractive(const Multilinear &poly, const std::vector<F> &challenges) { // :97-144
        SumcheckProof p;
        p.num_vars = poly.num_vars();
        p.round_polynomials.resize(p.num_vars ? p.num_vars : 1);
        p.final_point.resize(p.num_vars ? p.num_vars : 1);
        check(zh_sumcheck_prove_interactive(poly.context().get(), poly.handle(), challenges.data(), (uint32_t)challenges.size(),
                                            p.round_polynomials[0].data(), p.final_point.data(), &p.final_eval));
        p.round_polynomials.resize(p.num_vars);
        p.final_point.resize(p.num_vars);
        return p;
    }
};

// src/commitments/merkle_tree.zig:39-75
struct MerklePath {
    std::vector<Digest> siblings; // leaf -> root
    std::vector<bool> directions; // true = this node is the right child
};
struct MerkleOpeningProof {
    F value = 0;
    size_t index = 0;
    MerklePath path;
};

// src/commitments/merkle_tree.zig:273-402 with SHA3Hasher
class SimpleMerkleTree {
  public:
    static SimpleMerkleTree build(Context &c, const std::vector<F> &values) { // :283-318
        SimpleMerkleTree t(c);
        check(zb_merkle_build_values(c.get(), values.data(), values.size(), &t.h_, t.root_.data()));
        return t;
    }
    SimpleMerkleTree(SimpleMerkleTree &&o) noexcept : c_(o.c_), h_(o.h_), root_(o.root_) { o.h_ = 0; }
    ~SimpleMerkleTree() { deinit(); }
    void deinit() {
        if (h_) zb_merkle_free(c_->get(), h_);
        h_ = 0;
    }
    Digest getRoot() const { return root_; } // :320-322
    uint32_t height() const {
        uint32_t hgt = 0;
        check(zb_merkle_info(c_->get(), h_, nullptr, &hgt, nullptr));
        return hgt;
    }
    MerkleOpeningProof open(size_t index) const { // :324-360
        const uint32_t hgt = height();
        std::vector<uint8_t> sib((hgt ? hgt : 1) * 32), dirs(hgt ? hgt : 1);
        MerkleOpeningProof p;
        p.index = index;
        check(zb_merkle_open(c_->get(), h_, index, sib.data(), dirs.data(), &p.value));
        for (uint32_t l = 0; l < hgt; l++) {
            Digest d;
            std::copy(sib.begin() + 32 * l, sib.begin() + 32 * (l + 1), d.begin());
            p.path.siblings.push_back(d);
            p.path.directions.push_back(dirs[l] != 0);
        }
        return p;
    }
    static bool verify(const Digest &root, const MerkleOpeningProof &proof) { // :362-373
        const size_t hgt = proof.path.siblings.size();
        std::vector<uint8_t> sib((hgt ? hgt : 1) * 32), dirs(hgt ? hgt : 1);
        for (size_t l = 0; l < hgt; l++) {
            std::copy(proof.path.siblings[l].begin(), proof.path.siblings[l].end(), sib.begin() + 32 * l);
            dirs[l] = proof.path.directions[l] ? 1 : 0;
        }
        return zh_merkle_verify(root.data(), proof.value, sib.data(), dirs.data(), (uint32_t)hgt) != 0;
    }
    zb_tree handle() const { return h_; }

  private:
    explicit SimpleMerkleTree(Context &c) : c_(&c) {}
    Context *c_;
    zb_tree h_ = 0;
    Digest root_{};
    friend struct CommitmentScheme;
};

// src/commitments/polynomial_commit.zig:24-55
struct PolynomialCommitment {
    Digest commitment{};
    size_t num_vars = 0;
};
struct OpeningProof {
    std::vector<F> point;
    F value = 0;
    MerkleOpeningProof merkle_proof;
};

// src/commitments/polynomial_commit.zig:58-185 — CommitmentSchemeSHA3(BabyBear)
struct CommitmentScheme {
    static std::pair<PolynomialCommitment, SimpleMerkleTree> commit(const Multilinear &poly) { // :69-83
        SimpleMerkleTree t(poly.context());
        uint32_t v = 0;
        check(zh_commit(poly.context().get(), poly.handle(), &t.h_, t.root_.data(), &v));
        PolynomialCommitment c;
        c.commitment = t.root_;
        c.num_vars = v;
        return {c, std::move(t)};
    }
    static OpeningProof open(const Multilinear &poly, const SimpleMerkleTree &tree, const std::vector<F> &point) { // :86-115
        const uint32_t hgt = tree.height();
        std::vector<uint8_t> sib((hgt ? hgt : 1) * 32), dirs(hgt ? hgt : 1);
        OpeningProof p;
        p.point = point;
        uint64_t idx = 0;
        check(zh_commit_open(poly.context().get(), poly.handle(), tree.handle(), point.data(), (uint32_t)point.size(), &p.value,
                             &idx, &p.merkle_proof.value, sib.data(), dirs.data()));
        p.merkle_proof.index = idx;
        for (uint32_t l = 0; l < hgt; l++) {
            Digest d;
            std::copy(sib.begin() + 32 * l, sib.begin() + 32 * (l + 1), d.begin());
            p.merkle_proof.path.siblings.push_back(d);
            p.merkle_proof.path.directions.push_back(dirs[l] != 0);
        }
        return p;
    }
    static bool verify(const PolynomialCommitment &c, const OpeningProof &p) { // :118-129
        if (p.point.size() != c.num_vars) return false;
        return SimpleMerkleTree::verify(c.commitment, p.merkle_proof);
    }
};

// src/lookups/lasso_prover.zig:27-62
struct LassoProof {
    SumcheckProof sumcheck_proof;
    Digest query_commitment{}, table_commitment{};
    size_t num_lookups = 0;
};

// src/lookups/lasso_prover.zig:103-252; rows are flattened (inputs || outputs), `arity` values per row
struct LassoProver {
    static LassoProof prove(Context &c, const std::vector<F> &table_rows, const std::vector<F> &query_rows, uint32_t arity) {
        return run(c, table_rows, query_rows, nullptr, arity);
    }
    static LassoProof proveWithMapping(Context &c, const std::vector<F> &table_rows, const std::vector<F> &query_rows,
                                       const std::vector<uint64_t> &mapping, uint32_t arity) { // :179-205
        return run(c, table_rows, query_rows, &mapping, arity);
    }

  private:
    static LassoProof run(Context &c, const std::vector<F> &t, const std::vector<F> &q, const std::vector<uint64_t> *mapping,
                          uint32_t arity) {
        const uint64_t nt = t.size() / arity, nq = q.size() / arity;
        size_t vmax = 1;
        while ((1ull << vmax) < nq) vmax++;
        LassoProof p;
        p.num_lookups = nq;
        p.sumcheck_proof.round_polynomials.resize(vmax);
        p.sumcheck_proof.final_point.resize(vmax);
        uint32_t v = 0;
        const int32_t rc =
            mapping ? zh_lasso_prove_with_mapping(c.get(), t.data(), nt, q.data(), nq, mapping->data(), mapping->size(), arity,
                                                  p.sumcheck_proof.round_polynomials[0].data(), p.sumcheck_proof.final_point.data(),
                                                  &p.sumcheck_proof.final_eval, &v, p.query_commitment.data(), p.table_commitment.data())
                    : zh_lasso_prove(c.get(), t.data(), nt, q.data(), nq, arity, p.sumcheck_proof.round_polynomials[0].data(),
                                     p.sumcheck_proof.final_point.data(), &p.sumcheck_proof.final_eval, &v, p.query_commitment.data(),
                                     p.table_commitment.data());
        check(rc);
        p.sumcheck_proof.num_vars = v;
        p.sumcheck_proof.round_polynomials.resize(v);
        p.sumcheck_proof.final_point.resize(v);
        return p;
    }
};

} // namespace zigz
