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
// MAX_VAR_DEGREE covers the longest term.  Like the product kernels: warps take their work from a global chunk counter and
// the running products of all D+1 evaluation points are multiplied side by side (the deferred reduction of the last
// products exists as a variant, off by default: it costs more in occupancy than it saves here).  The launcher also
// factors terms that differ in one table — add.Wb + add.Wc becomes add.(Wb + Wc) through a "virtual table" that the
// kernel forms by one addition per item (sop_group below): the GKR layer costs 3 products per point instead of 4.
// Knobs for A/B runs: ZK_B200_SOP_WIDE=1 (deferred reduction), ZK_B200_SOP_FOLD_PIPE=f64 (FP64 folds),
// ZK_B200_SOP_GROUP=0 (no factoring), ZK_B200_SOP_SCHED=static.
#include <atomic>
#include <cstdlib>

#include "kernels.h"
#include "reduce.cuh"
#include "accw.cuh"
namespace zk {
namespace {
// hooks of sop_kernel.cuh's dynamic work distribution (the chunk counter sits next to the ticket, reduce.cuh)
__device__ __forceinline__ uint32_t sop_fetch_chunk(const ReduceArgs& ra) { return atomicAdd(ra.ticket + kWorkCounterOffset, 1u); }
__device__ __forceinline__ uint32_t sop_bcast_lane0(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
}  // namespace
}  // namespace zk
#include "sop_kernel.cuh"
#include "sop_group.hpp"

namespace zk {
namespace {

constexpr size_t sop_smem_bytes(int n_slots) { return (size_t)2 * n_slots * kThreads * sizeof(Fe); }  // real + virtual tables

constexpr size_t sop_smem_total(int n_slots, int np, bool wide) { return sop_smem_bytes(n_slots) + (wide ? accw_bytes(np) : 0); }

inline bool env_is(const char* name, char c) {
    const char* e = std::getenv(name);
    return e && e[0] == c;
}
// Which pipe folds: ZK_B200_SOP_FOLD_PIPE=int|f64 (default int: the FP64 variant measured 6 % slower on the GKR shape)
inline bool sop_fold_on_f64() {
    static const bool on = env_is("ZK_B200_SOP_FOLD_PIPE", 'f');
    return on;
}
// Deferred reduction of every term's last product: ZK_B200_SOP_WIDE=1.  Default off — measured on the GKR shape (4 x 2^24,
// profiles/r02_sop_variants.txt): its 34 KB of accumulators per block leave 3 resident blocks instead of 4 and the proof
// takes 5.19 ms against 4.73 ms with reduced products (the product kernels, with fewer tables in shared memory, gain from it).
inline bool sop_wide() {
    static const bool on = env_is("ZK_B200_SOP_WIDE", '1') && !sop_fold_on_f64();
    return on;
}
inline bool sop_dynamic() {
    static const bool on = !env_is("ZK_B200_SOP_SCHED", 's');
    return on;
}
// Toom point set (0, 1, -1, infinity) for MAX_VAR_DEGREE 3 when no term has more than three factors (ZK_B200_SOP_TOOM=0: plain 0..3)
inline bool sop_toom() {
    static const bool on = !env_is("ZK_B200_SOP_TOOM", '0');
    return on;
}
inline bool sop_grouping() {
    static const bool on = !env_is("ZK_B200_SOP_GROUP", '0');
    return on;
}

template <class F, int D, bool FOLD, bool F64, bool WIDE, bool DYN, bool TOOM = false>
cudaError_t do_sop_v(const TablePtrs& tabs, const SopSpec& spec, uint64_t q, const Fe& r, const ReduceScratch& s,
                     cudaStream_t st, const Fe* claim) {
    const int slots = spec.n_tables + spec.n_virt;
    const size_t smem = sop_smem_total(slots, D + 1, WIDE);
    static PerDeviceCache cache[kMaxFactors + kMaxVirtual + 1];  // per device and slot count (the shared-memory footprint depends on it)
    const int bpsm = per_device(cache[slots], [smem] {
        cudaError_t e = cudaFuncSetAttribute(sop_round_kernel<F, D, FOLD, F64, WIDE, DYN, TOOM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sop_smem_total(kMaxFactors + kMaxVirtual, D + 1, WIDE));
        if (e != cudaSuccess) return -(int)e;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sop_round_kernel<F, D, FOLD, F64, WIDE, DYN, TOOM>, kThreads, smem) != cudaSuccess || nb < 1) nb = 1;
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
    sop_round_kernel<F, D, FOLD, F64, WIDE, DYN, TOOM><<<grid, kThreads, smem, st>>>(tabs, spec, q, tab, tab64, ra);
    return cudaGetLastError();
}

template <class F, int D, bool FOLD>
cudaError_t do_sop(const TablePtrs& tabs, const SopSpec& spec_in, uint64_t q, const Fe& r, const ReduceScratch& s,
                   cudaStream_t st, const Fe* claim) {
    const SopSpec spec = sop_grouping() ? sop_group(spec_in) : spec_in;
    const bool dyn = sop_dynamic();
    if (D == 3 && sop_toom() && !sop_wide() && !(FOLD && sop_fold_on_f64()) && dyn) {  // the default configuration only
        bool cubic_at_most = true;
        for (int t = 0; t < spec.n_terms; t++) cubic_at_most &= spec.len[t] <= 3;
        if (cubic_at_most) return do_sop_v<F, D, FOLD, false, false, true, D == 3>(tabs, spec, q, r, s, st, claim);
    }
    if (sop_wide())
        return dyn ? do_sop_v<F, D, FOLD, false, true, true>(tabs, spec, q, r, s, st, claim) : do_sop_v<F, D, FOLD, false, true, false>(tabs, spec, q, r, s, st, claim);
    if (FOLD && sop_fold_on_f64())
        return dyn ? do_sop_v<F, D, FOLD, FOLD, false, true>(tabs, spec, q, r, s, st, claim) : do_sop_v<F, D, FOLD, FOLD, false, false>(tabs, spec, q, r, s, st, claim);
    return dyn ? do_sop_v<F, D, FOLD, false, false, true>(tabs, spec, q, r, s, st, claim) : do_sop_v<F, D, FOLD, false, false, false>(tabs, spec, q, r, s, st, claim);
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
    if (small_round_applies(spec.n_tables, degree, n_prev / 4))  // few items: the latency kernel (8 lanes per item)
        return launch_small_fold_round(field, tabs, spec, degree, n_prev / 4, r, scratch, stream, launches, claim);
    ++*launches;
    return field == Fr381::ID ? do_sop_deg<Fr381, true>(tabs, spec, degree, n_prev / 4, r, scratch, stream, claim)
                              : do_sop_deg<Fr377, true>(tabs, spec, degree, n_prev / 4, r, scratch, stream, claim);
}

}  // namespace zk
