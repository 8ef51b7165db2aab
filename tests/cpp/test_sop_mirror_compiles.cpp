// Compile-and-link check of the sum-of-products part of include/zk_b200.hpp.  Run without
// arguments it only touches host-side entry points (no GPU needed); the GPU parity of the same entry points is
// tests/test_gpu_sop.py.
#include <cstdio>
#include <string>

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

// `run`: the GKR layer end to end through the C++ mirror (needs a GPU — or the host mock of
// tests/test_hostmock_orchestration.py, which is how the CPU suite executes it)
static int run_gkr_layer() {
    const unsigned n = 6;
    std::vector<zk::MultiLinearPolynomial<F>> tabs;
    for (int k = 0; k < 4; k++) {
        std::vector<F> ev;
        for (unsigned i = 0; i < (1u << n); i++) ev.push_back(F(1000003ull * (k + 1) + 7919ull * i * i + i));
        tabs.emplace_back(n, ev);
    }
    // the same polynomial written the way a GKR prover would: add.(Wb + Wc) + mul.Wb.Wc, summed on the host
    F direct(0);
    {
        std::vector<std::vector<F>> e;
        for (auto& t : tabs) e.push_back(t.evaluation_slice());
        for (unsigned i = 0; i < (1u << n); i++) direct = direct + e[0][i] * (e[2][i] + e[3][i]) + e[1][i] * e[2][i] * e[3][i];
    }
    zk::SumOfProductsPoly<F> check_sum(std::move(tabs), {{0, 2}, {0, 3}, {1, 2, 3}});
    if (check_sum.sum() != direct) { std::printf("sum mismatch\n"); return 1; }
    auto keep = check_sum.clone();
    auto proof = zk::SumcheckProver<3, F>::prove(check_sum.clone(), direct);
    if (!zk::SumcheckVerifier<F>::verify(keep, proof)) { std::printf("verify rejected\n"); return 1; }
    auto partial = zk::SumcheckProver<3, F>::prove_partial(std::move(check_sum), direct);
    auto sub = zk::SumcheckVerifier<F>::verify_partial(partial.first);
    if (sub.challenges != partial.second || keep.evaluate(sub.challenges) != sub.sum) { std::printf("subclaim mismatch\n"); return 1; }
    std::printf("GKR LAYER OK\n");
    return 0;
}

int main(int argc, char** argv) {
    const uint8_t len[3] = {2, 2, 3}, fac[7] = {0, 2, 0, 3, 1, 2, 3};
    F v[4] = {F(3), F(5), F(7), F(11)}, out;
    zk::check(zk_sop_combine(F::field_id, len, fac, 3, v[0].limbs.data(), 4, out.limbs.data()));
    if (out != F(3 * (7 + 11) + 5 * 7 * 11)) { std::printf("zk_sop_combine mismatch\n"); return 1; }
    if (argc > 100) gkr_layer({});
    std::printf("SOP MIRROR OK\n");
    if (argc > 1 && std::string(argv[1]) == "run") return run_gkr_layer();
    return 0;
}
