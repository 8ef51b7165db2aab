// The reference's unit tests for the hot path, re-expressed on the C++ mirror (include/zk_b200.hpp).
// Built and run by tests/test_gpu_cpp_mirror.py on the GPU box.  Test names follow the reference.
#include <cstdio>
#include <cstring>
#include <string>

#include "zk_b200.hpp"

using Fr = zk::Fr381;
using MLP = zk::MultiLinearPolynomial<Fr>;
using PP = zk::ProductPoly<Fr>;
static int failures = 0;
#define EXPECT(cond)                                                         \
    do {                                                                     \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)
template <class Fn>
static std::string err_of(Fn&& f) {
    try { f(); } catch (const zk::Error& e) { return e.what(); }
    return "";
}
static std::vector<Fr> frs(std::initializer_list<long> v) {
    std::vector<Fr> r;
    for (long x : v) r.push_back(Fr::from_i64(x));
    return r;
}

// polynomial/src/multilinear/evaluation_form.rs:112-202
static void test_new_multilinear_poly() {
    EXPECT(err_of([] { MLP(2, frs({3, 1, 2})); }) == "evaluation vec len should equal 2^n_vars");
    EXPECT(err_of([] { MLP(2, frs({3, 1})); }) == "evaluation vec len should equal 2^n_vars");
    EXPECT(err_of([] { MLP(1, frs({3, 1})); }).empty());
    EXPECT(err_of([] { MLP(2, frs({3, 1, 2, 5})); }).empty());
}
static void test_partial_evaluate_single_variable() {
    MLP poly(2, frs({3, 1, 2, 5}));
    EXPECT(poly.partial_evaluate(0, frs({5})).evaluation_slice() == frs({-2, 21}));
    EXPECT(poly.partial_evaluate(0, frs({0})).evaluation_slice() == frs({3, 1}));
}
static void test_partial_evaluate_consecutive_variables() {
    MLP poly(3, frs({0, 0, 0, 3, 0, 0, 2, 5}));
    auto out = poly.partial_evaluate(1, frs({2, 3})).evaluation_slice();
    EXPECT(out.size() == 2 && out == frs({18, 22}));
}
static void test_full_evaluation() {
    MLP poly(3, frs({0, 0, 0, 3, 0, 0, 2, 5}));
    EXPECT(poly.evaluate(frs({2, 3, 4})) == Fr(48));
    EXPECT(err_of([&] { poly.evaluate(frs({2, 3})); }) == "evaluate must assign to all variables");
}
// polynomial/src/product_poly.rs:97-196
static void test_product_poly() {
    EXPECT(err_of([] { PP(std::vector<MLP>{}); }) == "cannot create product polynomial from empty polynomials");
    EXPECT(err_of([] { PP({MLP(1, frs({1, 2})), MLP(2, frs({1, 2, 3, 4}))}); }) ==
           "cannot create product polynomial from polynomial that don't share the same number of variables");
    MLP p1(3, frs({0, 0, 0, 3, 0, 0, 2, 5})), p2(3, frs({1, 2, 3, 4, 5, 6, 7, 8}));
    PP pp({p1, p2});
    auto pt = frs({2, 3, 4});
    EXPECT(pp.evaluate(pt) == p1.evaluate(pt) * p2.evaluate(pt));
    auto pe = pp.partial_evaluate(0, frs({7}));
    EXPECT(pe.polynomials()[0] == p1.partial_evaluate(0, frs({7})) && pe.polynomials()[1] == p2.partial_evaluate(0, frs({7})));
    EXPECT(PP({MLP(2, frs({2, 8, 10, 14})), MLP(2, frs({2, 8, 10, 22}))}).prod_reduce() == frs({4, 64, 100, 308}));
}
// sumcheck/src/lib.rs:53-122
static MLP p_2ab_3bc() { return MLP(3, frs({0, 0, 0, 3, 0, 0, 2, 5})); }
static void test_sumcheck_correct_sum_multilinear() {
    PP prod_poly({p_2ab_3bc()});
    auto proof = zk::SumcheckProver<1, Fr>::prove(prod_poly.clone(), Fr(10));
    EXPECT(zk::SumcheckVerifier<Fr>::verify(prod_poly, proof));
    // golden: round 0 = [3, 7]
    EXPECT(proof.round_polys.size() == 3 && proof.round_polys[0] == frs({3, 7}));
}
static void test_correct_sum_multivariate_deg_2() {
    PP p({MLP(2, frs({3, 3, 5, 5})), MLP(2, frs({0, 0, 0, 1}))});
    auto proof = zk::SumcheckProver<2, Fr>::prove(p.clone(), Fr(5));
    EXPECT(zk::SumcheckVerifier<Fr>::verify(p, proof));
    EXPECT(proof.round_polys[0] == frs({0, 5, 14}));
}
static void test_correct_sum_prove_partial() {
    PP prod_poly({p_2ab_3bc()});
    auto res = zk::SumcheckProver<1, Fr>::prove_partial(prod_poly.clone(), Fr(10));
    auto subclaim = zk::SumcheckVerifier<Fr>::verify_partial(res.first);
    EXPECT(prod_poly.evaluate(subclaim.challenges) == subclaim.sum);
    EXPECT(subclaim.challenges == res.second);
}
static void test_invalid_sum() {
    PP prod_poly({p_2ab_3bc()});
    auto proof = zk::SumcheckProver<1, Fr>::prove(prod_poly.clone(), Fr(12));
    EXPECT(err_of([&] { zk::SumcheckVerifier<Fr>::verify(prod_poly, proof); }) == "verifier check failed: claimed_sum != p(0) + p(1)");
}
// fft/src/lib.rs:78-82
static void test_fft() {
    using Fq = zk::Fr377;
    std::vector<Fq> a = {Fq(0), Fq(2), Fq(34), Fq(3434)};
    EXPECT(zk::ifft(zk::fft(a)) == a);
    EXPECT(err_of([] { zk::fft(std::vector<Fq>{Fq(1), Fq(2), Fq(3)}); }) == "values must be a power of 2");
}

int main() {
    test_new_multilinear_poly();
    test_partial_evaluate_single_variable();
    test_partial_evaluate_consecutive_variables();
    test_full_evaluation();
    test_product_poly();
    test_sumcheck_correct_sum_multilinear();
    test_correct_sum_multivariate_deg_2();
    test_correct_sum_prove_partial();
    test_invalid_sum();
    test_fft();
    std::printf(failures ? "FAILED (%d)\n" : "ALL C++ MIRROR TESTS PASSED\n", failures);
    return failures ? 1 : 0;
}
