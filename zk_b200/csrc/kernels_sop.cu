// kernels_sop.cu — sumcheck rounds over a SUM of products of MLE tables on sm_100a.
//
//   P(x) = sum_t prod_{k in term t} A_k(x)        e.g. the GKR layer polynomial add.Wb + add.Wc + mul.Wb.Wc
//
// The reference has no such type: its ProductPoly is one product (polynomial/src/product_poly.rs:4-10) and readme.md:9
// only points at a GKR crate that is not in the tree.  The protocol is the reference's own (sumcheck/src/prover.rs:33-73):
// per round the evaluations S(t), t = 0..D, of the sum over the remaining hypercube with variable 0 bound to t, then
// every table is folded at the challenge (evaluation_form.rs:40-80, pair (j, j + N/2), l - r (l - h)).  A single-term
// sum is exactly the reference's ProductPoly proof (tests pin that), and S is linear in the terms.
//
// One pass per round, like the product kernels: round 0 reads every table once; every later round folds the
// quadruple (j, j+q, j+2q, j+3q) of every table at the previous challenge, writes the two folded values back in
// place and forms the next round's sums from them — a table shared by several terms is read and folded once.
// The per-item values e_k(t) = lo_k + t (hi_k - lo_k) of all tables live in shared memory (2 x n_tables elements per
// thread) because the terms index them with run-time table numbers; the products run on the general fe_mul.
// Rounds >= 1 derive S(1) from the previous round polynomial (one evaluation point less to multiply out) when
// MAX_VAR_DEGREE covers the longest term; FP64 folds are wired behind ZK_B200_SOP_FOLD_PIPE.  No deferred
// reduction or dynamic chunks yet (kernels_sumcheck.cu has those).
#include <atomic>
#include <cstdlib>

#include "kernels.h"
#include "reduce.cuh"
#include "accw.cuh"
#include "sop_kernel.cuh"

namespace zk {
namespace {

constexpr size_t sop_smem_bytes(int n_tables) { return (size_t)2 * n_tables * kThreads * sizeof(Fe); }

constexpr size_t sop_smem_total(int n_tables, int np, bool wide) { return sop_smem_bytes(n_tables) + (wide ? accw_bytes(np) : 0); }

template <class F, int D, bool FOLD, bool F64, bool WIDE>
cudaError_t do_sop_v(const TablePtrs& tabs, const SopSpec& spec, uint64_t q, const Fe& r, const ReduceScratch& s,
                     cudaStream_t st, const Fe* claim) {
    const size_t smem = sop_smem_total(spec.n_tables, D + 1, WIDE);
    static PerDeviceCache cache[kMaxFactors + 1];  // per device and table count (the shared-memory footprint depends on it)
    const int bpsm = per_device(cache[spec.n_tables], [smem] {
        cudaError_t e = cudaFuncSetAttribute(sop_round_kernel<F, D, FOLD, F64, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sop_smem_total(kMaxFactors, D + 1, WIDE));
        if (e != cudaSuccess) return -(int)e;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sop_round_kernel<F, D, FOLD, F64, WIDE>, kThreads, smem) != cudaSuccess || nb < 1) nb = 1;
        return nb;
    });
    if (bpsm <= 0) return (cudaError_t)(-bpsm);
    const unsigned grid = grid_for(q, kThreads, s.num_sms, bpsm);
    const FixedMul tab = (FOLD && !F64) ? make_fixed<F>(r) : FixedMul{};
    const FixedMulF64Sel tab64 = F64 ? make_fixed_f64<F>(r) : FixedMulF64Sel{};
    ReduceArgs ra = make_ra(s, 0);
    if (FOLD && claim) {
        ra.skip1 = 1;
        ra.claim = *claim;
    }
    sop_round_kernel<F, D, FOLD, F64, WIDE><<<grid, kThreads, smem, st>>>(tabs, spec, q, tab, tab64, ra);
    return cudaGetLastError();
}

// Which pipe folds: ZK_B200_SOP_FOLD_PIPE=int|f64 (default int: the FP64 variant measured 6 % slower on the GKR shape)
inline bool sop_fold_on_f64() {
    static const bool on = [] {
        const char* e = std::getenv("ZK_B200_SOP_FOLD_PIPE");
        return e && e[0] == 'f';
    }();
    return on;
}
// ZK_B200_SOP_WIDE=1: deferred reduction of every term's last product (default off: written after the round's GPU
// budget was spent — replayed on the host, not yet measured or run on hardware)
inline bool sop_wide() {
    static const bool on = [] {
        const char* e = std::getenv("ZK_B200_SOP_WIDE");
        return e && e[0] == '1';
    }();
    return on;
}

template <class F, int D, bool FOLD>
cudaError_t do_sop(const TablePtrs& tabs, const SopSpec& spec, uint64_t q, const Fe& r, const ReduceScratch& s,
                   cudaStream_t st, const Fe* claim) {
    if (sop_wide()) return do_sop_v<F, D, FOLD, false, true>(tabs, spec, q, r, s, st, claim);
    if (FOLD && sop_fold_on_f64()) return do_sop_v<F, D, FOLD, FOLD, false>(tabs, spec, q, r, s, st, claim);
    return do_sop_v<F, D, FOLD, false, false>(tabs, spec, q, r, s, st, claim);
}

template <class F, bool FOLD>
cudaError_t do_sop_deg(const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t q, const Fe& r,
                       const ReduceScratch& s, cudaStream_t st, const Fe* claim = nullptr) {
    switch (degree) {
        case 1: return do_sop<F, 1, FOLD>(tabs, spec, q, r, s, st, claim);
        case 2: return do_sop<F, 2, FOLD>(tabs, spec, q, r, s, st, claim);
        case 3: return do_sop<F, 3, FOLD>(tabs, spec, q, r, s, st, claim);
        case 4: return do_sop<F, 4, FOLD>(tabs, spec, q, r, s, st, claim);
        default: return cudaErrorInvalidValue;
    }
}

bool spec_ok(const SopSpec& spec) {
    if (spec.n_tables < 1 || spec.n_tables > kMaxFactors || spec.n_terms < 1 || spec.n_terms > kMaxTerms) return false;
    for (int t = 0; t < spec.n_terms; t++) {
        if (spec.len[t] < 1 || spec.len[t] > kMaxFactors) return false;
        for (int i = 0; i < (int)spec.len[t]; i++)
            if ((int)spec.fac[t][i] >= spec.n_tables) return false;
    }
    return true;
}

}  // namespace

bool sop_degree_supported(int degree) { return degree >= 1 && degree <= 4; }

cudaError_t launch_sop_round_poly(int field, const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t half,
                                  const ReduceScratch& scratch, cudaStream_t stream, int* launches) {
    if (!spec_ok(spec) || half < 1) return cudaErrorInvalidValue;
    ++*launches;
    return field == Fr381::ID ? do_sop_deg<Fr381, false>(tabs, spec, degree, half, Fe{}, scratch, stream)
                              : do_sop_deg<Fr377, false>(tabs, spec, degree, half, Fe{}, scratch, stream);
}

cudaError_t launch_sop_fold_round_poly(int field, const TablePtrs& tabs, const SopSpec& spec, int degree,
                                       uint64_t n_prev, const Fe& r, const ReduceScratch& scratch, cudaStream_t stream,
                                       int* launches, const Fe* claim) {
    if (!spec_ok(spec) || n_prev < 4) return cudaErrorInvalidValue;
    ++*launches;
    return field == Fr381::ID ? do_sop_deg<Fr381, true>(tabs, spec, degree, n_prev / 4, r, scratch, stream, claim)
                              : do_sop_deg<Fr377, true>(tabs, spec, degree, n_prev / 4, r, scratch, stream, claim);
}

}  // namespace zk
