// zk_b200.hpp — C++ host-side mirror of the reference's Rust API over the C ABI (zk_b200.h).
//
// The reference is compiled code (Rust); with no Rust toolchain in the build image this header is the
// compiled-language host layer: same type and method names, argument meaning and error behaviour as
//   polynomial::multilinear::evaluation_form::MultiLinearPolynomial   (evaluation_form.rs:7-103)
//   polynomial::product_poly::ProductPoly                              (product_poly.rs:4-88)
//   sumcheck::{SumcheckProof, SubClaim, prover::SumcheckProver, verifier::SumcheckVerifier}
//   transcript::Transcript                                             (transcript/src/lib.rs:5-35)
//   fft::{fft, ifft}                                                   (fft/src/lib.rs:4-19)
// `Result<_, &'static str>` becomes a thrown zk::Error whose what() is the reference's literal string.
// Header only; link with -lzk_b200.  All arithmetic runs in the CUDA library (no CPU fallback).
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "zk_b200.h"

namespace zk {

struct Error : std::runtime_error {
    int status;
    Error(int st) : std::runtime_error(zk_status_string(st)), status(st) {}
};
inline void check(int st) {
    if (st != ZK_OK) throw Error(st);
}

// A field element exactly as ark-ff keeps it: 4 LE u64 limbs, Montgomery form.
template <int FIELD>
struct Fp {
    std::array<uint64_t, 4> limbs{};
    static constexpr int field_id = FIELD;
    Fp() = default;
    Fp(uint64_t x) { check(zk_field_from_u64(FIELD, x, limbs.data())); }  // F::from(u64)
    static Fp from_i64(int64_t x) { return x >= 0 ? Fp((uint64_t)x) : Fp(0) - Fp((uint64_t)(-x)); }
    friend Fp operator+(const Fp& a, const Fp& b) { Fp r; zk_field_add(FIELD, a.limbs.data(), b.limbs.data(), r.limbs.data()); return r; }
    friend Fp operator-(const Fp& a, const Fp& b) { Fp r; zk_field_sub(FIELD, a.limbs.data(), b.limbs.data(), r.limbs.data()); return r; }
    friend Fp operator*(const Fp& a, const Fp& b) { Fp r; zk_field_mul(FIELD, a.limbs.data(), b.limbs.data(), r.limbs.data()); return r; }
    friend bool operator==(const Fp& a, const Fp& b) { return a.limbs == b.limbs; }
    friend bool operator!=(const Fp& a, const Fp& b) { return !(a == b); }
};
using Fr381 = Fp<ZK_BLS12_381_FR>;  // ark_bls12_381::Fr
using Fr377 = Fp<ZK_BLS12_377_FR>;  // ark_bls12_377::Fr

class Context {
   public:
    explicit Context(int device = 0) { check(zk_ctx_create(device, &c_)); }
    ~Context() { zk_ctx_destroy(c_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    zk_ctx* get() const { return c_; }
    static Context& instance() {
        static Context ctx(0);
        return ctx;
    }

   private:
    zk_ctx* c_ = nullptr;
};

template <class F>
class MultiLinearPolynomial {
   public:
    // new(n_vars, evaluations) -> Result  (evaluation_form.rs:15-27)
    MultiLinearPolynomial(size_t n_vars, const std::vector<F>& evaluations) {
        zk_table* t = nullptr;
        static const uint64_t dummy[4] = {0, 0, 0, 0};
        const uint64_t* p = evaluations.empty() ? dummy : evaluations[0].limbs.data();
        check(zk_table_upload(Context::instance().get(), F::field_id, p, evaluations.size(), (unsigned)n_vars, &t));
        t_.reset(t, zk_table_free);
    }
    size_t n_vars() const { return zk_table_n_vars(t_.get()); }  // :30
    // partial_evaluate(&self, initial_var, assignments) -> Result<Self>  (:40-80)
    MultiLinearPolynomial partial_evaluate(size_t initial_var, const std::vector<F>& assignments) const {
        zk_table* out = nullptr;
        check(zk_mle_partial_evaluate(Context::instance().get(), t_.get(), (unsigned)initial_var,
                                      assignments.empty() ? nullptr : assignments[0].limbs.data(), (unsigned)assignments.size(), &out));
        return MultiLinearPolynomial(out);
    }
    F evaluate(const std::vector<F>& assignments) const {  // :83-89
        F r;
        check(zk_mle_evaluate(Context::instance().get(), t_.get(), assignments.empty() ? nullptr : assignments[0].limbs.data(),
                              (unsigned)assignments.size(), r.limbs.data()));
        return r;
    }
    std::vector<F> evaluation_slice() const {  // :92-94 (a copy: the table lives on the device)
        std::vector<F> v(zk_table_local_len(t_.get()));
        if (!v.empty()) check(zk_table_download(Context::instance().get(), t_.get(), v[0].limbs.data()));
        return v;
    }
    std::vector<uint8_t> to_bytes() const {  // :97-103
        std::vector<uint8_t> b(32 * zk_table_local_len(t_.get()));
        if (!b.empty()) check(zk_mle_to_bytes(Context::instance().get(), t_.get(), b.data()));
        return b;
    }
    MultiLinearPolynomial clone() const {
        zk_table* out = nullptr;
        check(zk_table_clone(Context::instance().get(), t_.get(), &out));
        return MultiLinearPolynomial(out);
    }
    bool operator==(const MultiLinearPolynomial& o) const { return n_vars() == o.n_vars() && evaluation_slice() == o.evaluation_slice(); }
    zk_table* handle() const { return t_.get(); }

   private:
    explicit MultiLinearPolynomial(zk_table* t) : t_(t, zk_table_free) {}
    std::shared_ptr<zk_table> t_;
    template <class G>
    friend class ProductPoly;
};

template <class F>
class ProductPoly {
   public:
    // new(polynomials) -> Result  (product_poly.rs:14-32)
    explicit ProductPoly(std::vector<MultiLinearPolynomial<F>> polynomials) : polys_(std::move(polynomials)) {
        auto h = handles();
        check(zk_product_check(h.empty() ? nullptr : (const zk_table* const*)h.data(), (unsigned)h.size()));
    }
    size_t n_vars() const { return polys_[0].n_vars(); }  // :86
    F evaluate(const std::vector<F>& a) const {            // :36-44
        F r;
        auto h = handles();
        check(zk_product_evaluate(Context::instance().get(), (const zk_table* const*)h.data(), (unsigned)h.size(),
                                  a.empty() ? nullptr : a[0].limbs.data(), (unsigned)a.size(), r.limbs.data()));
        return r;
    }
    ProductPoly partial_evaluate(size_t initial_var, const std::vector<F>& a) const {  // :48-63
        std::vector<MultiLinearPolynomial<F>> out;
        for (auto& p : polys_) out.push_back(p.partial_evaluate(initial_var, a));
        return ProductPoly(std::move(out));
    }
    std::vector<F> prod_reduce() const {  // :66-74
        zk_table* out = nullptr;
        auto h = handles();
        check(zk_product_prod_reduce(Context::instance().get(), (const zk_table* const*)h.data(), (unsigned)h.size(), &out));
        return MultiLinearPolynomial<F>(out).evaluation_slice();
    }
    std::vector<uint8_t> to_bytes() const {  // :77-83
        std::vector<uint8_t> b;
        for (auto& p : polys_) { auto x = p.to_bytes(); b.insert(b.end(), x.begin(), x.end()); }
        return b;
    }
    ProductPoly clone() const {
        std::vector<MultiLinearPolynomial<F>> out;
        for (auto& p : polys_) out.push_back(p.clone());
        return ProductPoly(std::move(out));
    }
    std::vector<zk_table*> handles() const {
        std::vector<zk_table*> h;
        for (auto& p : polys_) h.push_back(p.handle());
        return h;
    }
    const std::vector<MultiLinearPolynomial<F>>& polynomials() const { return polys_; }

   private:
    std::vector<MultiLinearPolynomial<F>> polys_;
};

// P(x) = sum_t prod_{k in terms[t]} polynomials[k](x)  (beyond the reference's ProductPoly): the GKR
// layer polynomial add.Wb + add.Wc + mul.Wb.Wc is polynomials = {add, mul, Wb, Wc}, terms = {{0,2},{0,3},{1,2,3}}.
// Same surface as ProductPoly; a single term listing every table once is the reference's ProductPoly.
template <class F>
class SumOfProductsPoly {
   public:
    SumOfProductsPoly(std::vector<MultiLinearPolynomial<F>> polynomials, std::vector<std::vector<uint8_t>> terms)
        : polys_(std::move(polynomials)), terms_(std::move(terms)) {
        auto h = handles();
        check(zk_product_check(h.empty() ? nullptr : (const zk_table* const*)h.data(), (unsigned)h.size()));
        if (terms_.empty()) throw Error(ZK_ERR_EMPTY_PRODUCT);
        for (auto& t : terms_) {
            if (t.empty()) throw Error(ZK_ERR_EMPTY_PRODUCT);
            for (uint8_t k : t)
                if (k >= polys_.size()) throw Error(ZK_ERR_INVALID_ARG);
            len_.push_back((uint8_t)t.size());
            fac_.insert(fac_.end(), t.begin(), t.end());
        }
    }
    size_t n_vars() const { return polys_[0].n_vars(); }
    F sum() const {  // the honest claim
        F r;
        auto h = handles();
        check(zk_sop_sum(Context::instance().get(), (const zk_table* const*)h.data(), (unsigned)h.size(), len_.data(), fac_.data(),
                         (unsigned)len_.size(), r.limbs.data()));
        return r;
    }
    F evaluate(const std::vector<F>& a) const {
        F r;
        auto h = handles();
        check(zk_sop_evaluate(Context::instance().get(), (const zk_table* const*)h.data(), (unsigned)h.size(), len_.data(), fac_.data(),
                              (unsigned)len_.size(), a.empty() ? nullptr : a[0].limbs.data(), (unsigned)a.size(), r.limbs.data()));
        return r;
    }
    std::vector<F> round_poly(unsigned degree) const {  // prover.rs:48-56 for the sum of products
        std::vector<F> out(degree + 1);
        auto h = handles();
        check(zk_sop_round_poly(Context::instance().get(), (const zk_table* const*)h.data(), (unsigned)h.size(), len_.data(), fac_.data(),
                                (unsigned)len_.size(), degree, out[0].limbs.data()));
        return out;
    }
    SumOfProductsPoly partial_evaluate(size_t initial_var, const std::vector<F>& a) const {
        std::vector<MultiLinearPolynomial<F>> out;
        for (auto& p : polys_) out.push_back(p.partial_evaluate(initial_var, a));
        return SumOfProductsPoly(std::move(out), terms_);
    }
    std::vector<uint8_t> to_bytes() const {
        std::vector<uint8_t> b;
        for (auto& p : polys_) { auto x = p.to_bytes(); b.insert(b.end(), x.begin(), x.end()); }
        return b;
    }
    SumOfProductsPoly clone() const {
        std::vector<MultiLinearPolynomial<F>> out;
        for (auto& p : polys_) out.push_back(p.clone());
        return SumOfProductsPoly(std::move(out), terms_);
    }
    std::vector<zk_table*> handles() const {
        std::vector<zk_table*> h;
        for (auto& p : polys_) h.push_back(p.handle());
        return h;
    }
    const std::vector<uint8_t>& term_len() const { return len_; }
    const std::vector<uint8_t>& term_factors() const { return fac_; }

   private:
    std::vector<MultiLinearPolynomial<F>> polys_;
    std::vector<std::vector<uint8_t>> terms_;
    std::vector<uint8_t> len_, fac_;
};

template <class F>
struct SumcheckProof {  // sumcheck/src/lib.rs:8-11
    F sum;
    std::vector<std::vector<F>> round_polys;
};
template <class F>
struct SubClaim {  // sumcheck/src/lib.rs:17-20
    F sum;
    std::vector<F> challenges;
};

template <uint8_t MAX_VAR_DEGREE, class F>
struct SumcheckProver {  // sumcheck/src/prover.rs:9-74
    // prove(poly, sum): takes `poly` by value like the reference (the device tables are consumed)
    static SumcheckProof<F> prove(ProductPoly<F> poly, const F& sum) { return run(std::move(poly), sum, true).first; }
    static std::pair<SumcheckProof<F>, std::vector<F>> prove_partial(ProductPoly<F> poly, const F& sum) {
        return run(std::move(poly), sum, false);
    }

    // the same loop over a sum of products (zk_sumcheck_prove_sop)
    static SumcheckProof<F> prove(SumOfProductsPoly<F> poly, const F& sum) { return run_sop(std::move(poly), sum, true).first; }
    static std::pair<SumcheckProof<F>, std::vector<F>> prove_partial(SumOfProductsPoly<F> poly, const F& sum) {
        return run_sop(std::move(poly), sum, false);
    }

   private:
    static std::pair<SumcheckProof<F>, std::vector<F>> run_sop(SumOfProductsPoly<F> poly, const F& sum, bool absorb) {
        const size_t n = poly.n_vars(), np = (size_t)MAX_VAR_DEGREE + 1;
        std::vector<F> rp(n * np + 1), ch(n + 1);
        auto h = poly.handles();
        check(zk_sumcheck_prove_sop(Context::instance().get(), h.data(), (unsigned)h.size(), poly.term_len().data(),
                                    poly.term_factors().data(), (unsigned)poly.term_len().size(), MAX_VAR_DEGREE, sum.limbs.data(),
                                    absorb ? 1 : 0, rp[0].limbs.data(), ch[0].limbs.data(), nullptr));
        SumcheckProof<F> proof{sum, {}};
        for (size_t i = 0; i < n; i++) proof.round_polys.emplace_back(rp.begin() + i * np, rp.begin() + (i + 1) * np);
        ch.resize(n);
        return {proof, ch};
    }
    static std::pair<SumcheckProof<F>, std::vector<F>> run(ProductPoly<F> poly, const F& sum, bool absorb) {
        const size_t n = poly.n_vars(), np = (size_t)MAX_VAR_DEGREE + 1;
        std::vector<F> rp(n * np + 1), ch(n + 1);
        auto h = poly.handles();
        check(zk_sumcheck_prove(Context::instance().get(), h.data(), (unsigned)h.size(), MAX_VAR_DEGREE, sum.limbs.data(), absorb ? 1 : 0,
                                rp[0].limbs.data(), ch[0].limbs.data(), nullptr));
        SumcheckProof<F> proof{sum, {}};
        for (size_t i = 0; i < n; i++) proof.round_polys.emplace_back(rp.begin() + i * np, rp.begin() + (i + 1) * np);
        ch.resize(n);
        return {proof, ch};
    }
};

template <class F>
struct SumcheckVerifier {  // sumcheck/src/verifier.rs:9-79
    static bool verify(const ProductPoly<F>& poly, const SumcheckProof<F>& proof) {  // :15-33
        std::vector<F> flat;
        size_t np = proof.round_polys.empty() ? 1 : proof.round_polys[0].size();
        for (auto& r : proof.round_polys) flat.insert(flat.end(), r.begin(), r.end());
        auto h = poly.handles();
        int st = zk_sumcheck_verify(Context::instance().get(), (const zk_table* const*)h.data(), (unsigned)h.size(), proof.sum.limbs.data(),
                                    flat.empty() ? nullptr : flat[0].limbs.data(), (unsigned)proof.round_polys.size(), (unsigned)np - 1);
        if (st == ZK_VERIFY_FALSE) return false;  // Ok(false)
        check(st);
        return true;
    }
    // the same for a sum of products proved with SumcheckProver::prove (zk_sumcheck_verify_sop)
    static bool verify(const SumOfProductsPoly<F>& poly, const SumcheckProof<F>& proof) {
        std::vector<F> flat;
        size_t np = proof.round_polys.empty() ? 1 : proof.round_polys[0].size();
        for (auto& r : proof.round_polys) flat.insert(flat.end(), r.begin(), r.end());
        auto h = poly.handles();
        int st = zk_sumcheck_verify_sop(Context::instance().get(), (const zk_table* const*)h.data(), (unsigned)h.size(), poly.term_len().data(),
                                        poly.term_factors().data(), (unsigned)poly.term_len().size(), proof.sum.limbs.data(),
                                        flat.empty() ? nullptr : flat[0].limbs.data(), (unsigned)proof.round_polys.size(), (unsigned)np - 1);
        if (st == ZK_VERIFY_FALSE) return false;
        check(st);
        return true;
    }
    static SubClaim<F> verify_partial(const SumcheckProof<F>& proof) {  // :38-41
        std::vector<F> flat;
        size_t np = proof.round_polys.empty() ? 1 : proof.round_polys[0].size();
        for (auto& r : proof.round_polys) flat.insert(flat.end(), r.begin(), r.end());
        SubClaim<F> sc;
        sc.challenges.resize(proof.round_polys.size() + 1);
        check(zk_sumcheck_verify_partial(F::field_id, proof.sum.limbs.data(), flat.empty() ? nullptr : flat[0].limbs.data(),
                                         (unsigned)proof.round_polys.size(), (unsigned)np - 1, sc.sum.limbs.data(), sc.challenges[0].limbs.data()));
        sc.challenges.resize(proof.round_polys.size());
        return sc;
    }
};

class Transcript {  // transcript/src/lib.rs:5-35
   public:
    Transcript() : t_(zk_transcript_new()) {}
    ~Transcript() { zk_transcript_free(t_); }
    Transcript(const Transcript&) = delete;
    void append(const std::vector<uint8_t>& d) { zk_transcript_append(t_, d.data(), d.size()); }
    template <class F>
    F sample_field_element() {
        F r;
        check(zk_transcript_sample_field_element(t_, F::field_id, r.limbs.data()));
        return r;
    }
    template <class F>
    std::vector<F> sample_n_field_elements(size_t n) {
        std::vector<F> v;
        for (size_t i = 0; i < n; i++) v.push_back(sample_field_element<F>());
        return v;
    }

   private:
    zk_transcript* t_;
};

// fft/src/lib.rs:4-19.  The reference panics on bad lengths; here those are zk::Error 9 / 10.
template <class F>
std::vector<F> fft(std::vector<F> coefficients) {
    check(zk_ntt_host(Context::instance().get(), F::field_id, coefficients.empty() ? nullptr : coefficients[0].limbs.data(), coefficients.size(), 0));
    return coefficients;
}
template <class F>
std::vector<F> ifft(std::vector<F> evaluations) {
    check(zk_ntt_host(Context::instance().get(), F::field_id, evaluations.empty() ? nullptr : evaluations[0].limbs.data(), evaluations.size(), 1));
    return evaluations;
}

}  // namespace zk
