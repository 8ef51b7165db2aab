// Compile-and-link check of the sum-of-products part of include/zk_b200.hpp.  Run without
// arguments it only touches host-side entry points (no GPU needed); the GPU parity of the same entry points is
// tests/test_gpu_sop.py.
#include <cstdio>

#include "zk_b200.hpp"

using F = zk::Fr381;

// instantiated, never called without a GPU
std::pair<zk::SumcheckProof<F>, std::vector<F>> gkr_layer(std::vector<zk::MultiLinearPolynomial<F>> tabs) {
    zk::SumOfProductsPoly<F> p(std::move(tabs), {{0, 2}, {0, 3}, {1, 2, 3}});
    const F claim = p.sum();
    auto rp = p.round_poly(3);
    (void)rp;
    auto q = p.clone();
    auto proof = zk::SumcheckProver<3, F>::prove_partial(std::move(p), claim);
    auto sub = zk::SumcheckVerifier<F>::verify_partial(proof.first);
    if (q.evaluate(sub.challenges) != sub.sum) throw zk::Error(ZK_VERIFY_FALSE);
    return proof;
}

int main(int argc, char**) {
    const uint8_t len[3] = {2, 2, 3}, fac[7] = {0, 2, 0, 3, 1, 2, 3};
    F v[4] = {F(3), F(5), F(7), F(11)}, out;
    zk::check(zk_sop_combine(F::field_id, len, fac, 3, v[0].limbs.data(), 4, out.limbs.data()));
    if (out != F(3 * (7 + 11) + 5 * 7 * 11)) { std::printf("zk_sop_combine mismatch\n"); return 1; }
    if (argc > 100) gkr_layer({});
    std::printf("SOP MIRROR OK\n");
    return 0;
}
