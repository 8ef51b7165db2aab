// sop_kernel.cuh — the sum-of-products round kernel (see kernels_sop.cu for what it computes and why).
// Kept in its own header so that tests/cpp/test_sop_kernel_host.cpp can replay the very same source on the host,
// thread by thread, with host_field.hpp standing in for the device arithmetic: the index conventions (pairs, the
// in-place quadruple fold, term bookkeeping) are then checked on the CPU, without a GPU.
// Needs in scope: Fe / FixedMul / TablePtrs / SopSpec (kernels.h), FixedMulF64Sel, the fe_* and ld/st functions
// (field.cuh, field_f64.cuh on the device), kThreads, ReduceArgs and reduce_publish (reduce.cuh on the device), the
// Accw accumulators (accw.cuh on the device).
#pragma once

namespace zk {
namespace {

// F64 (only with FOLD): the folds run on the FP64 pipe (field_f64.cuh) instead of fe_mul_fixed's wide multiplies.
// ra.skip1 (only with FOLD): S(0) + S(1) of this round is known to the host (the previous round polynomial at its
// challenge), so the t = 1 products are skipped and the last block publishes S(1) = claim - S(0) (reduce_publish).
// WIDE: the LAST multiplication of every term is a plain 512-bit product (fe_mul_wide, 64 instead of 112 wide
// multiplies) added to a 17-word per-thread accumulator per evaluation point in shared memory (accw.cuh, as in the
// product kernels), Montgomery-reduced once per thread at the end; single-factor terms add x * 2^256.  Costs
// accw_bytes(D+1) more shared memory per block.  Same field elements (sum of products then one REDC == sum of REDCs).
template <class F, int D, bool FOLD, bool F64 = false, bool WIDE = false>
__global__ void __launch_bounds__(kThreads)
    sop_round_kernel(TablePtrs tabs, const __grid_constant__ SopSpec spec, uint64_t q,
                     const __grid_constant__ FixedMul rtab, const __grid_constant__ FixedMulF64Sel rtab64, ReduceArgs ra) {
    static_assert(!F64 || FOLD, "the FP64 variant only changes the folds");
    const bool skip1 = FOLD && ra.skip1 != 0;
    extern __shared__ __align__(32) uint4 sop_smem[];  // Fe [2 * n_tables][kThreads]: e_k then d_k; WIDE: + accumulators
    Fe* const ev = reinterpret_cast<Fe*>(sop_smem) + threadIdx.x;
    const int nt = spec.n_tables;
    Fe* const dv = ev + (size_t)nt * kThreads;
    Accw accw{};
    if (WIDE) {
        uint4* const accw_all = sop_smem + (size_t)2 * nt * kThreads * (sizeof(Fe) / sizeof(uint4));
        accw_zero(accw_all, D + 1);
        __syncthreads();
        accw = accw_base(accw_all, D + 1);
    }
    Fe acc[D + 1];
#pragma unroll
    for (int t = 0; t <= D; t++) acc[t] = fe_zero<F>();
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
#pragma unroll 1
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < q; j += stride) {
#pragma unroll 1
        for (int k = 0; k < nt; k++) {
            Fe* T = tabs.t[k];
            Fe lo, hi;
            if (FOLD) {  // T has 4q entries: fold (j, j+2q) and (j+q, j+3q) at r, write back to j and j+q
                const Fe x0 = ld_fe_stream(T + j), x2 = ld_fe_stream(T + j + 2 * q);
                const Fe x1 = ld_fe_stream(T + j + q), x3 = ld_fe_stream(T + j + 3 * q);
                if (F64) {
                    // `k >> 16` is always 0, but loop-variant for ptxas: see FixedMulF64Sel
                    fe_fold_fixed_f64_x2<F>(lo, hi, x0, x1, x2, x3, rtab64.t[k >> 16]);
                } else {
                    lo = fe_fold_fixed<F>(x0, x2, rtab);
                    hi = fe_fold_fixed<F>(x1, x3, rtab);
                }
                st_fe(T + j, lo);
                st_fe(T + j + q, hi);
            } else {  // T has 2q entries: the pair is (j, j+q)
                lo = ld_fe_stream(T + j);
                hi = ld_fe_stream(T + j + q);
            }
            ev[(size_t)k * kThreads] = lo;
            dv[(size_t)k * kThreads] = fe_sub<F>(hi, lo);
        }
#pragma unroll
        for (int t = 0; t <= D; t++) {
            if (!(t == 1 && skip1)) {
                if (WIDE) {
                    const Accw at = accw_at(accw, t);
#pragma unroll 1
                    for (int term = 0; term < spec.n_terms; term++) {
                        const int len = (int)spec.len[term];
                        Fe p = ev[(size_t)spec.fac[term][0] * kThreads];
                        if (len == 1) {
                            accw_add_hi(at, p);
                        } else {
#pragma unroll 1
                            for (int i = 1; i < len - 1; i++) p = fe_mul<F>(p, ev[(size_t)spec.fac[term][i] * kThreads]);
                            uint32_t w[16];
                            fe_mul_wide(w, p, ev[(size_t)spec.fac[term][len - 1] * kThreads]);
                            accw_add16(at, w);
                        }
                    }
                } else {
                    Fe s = fe_zero<F>();
#pragma unroll 1
                    for (int term = 0; term < spec.n_terms; term++) {
                        Fe p = ev[(size_t)spec.fac[term][0] * kThreads];
#pragma unroll 1
                        for (int i = 1; i < (int)spec.len[term]; i++) p = fe_mul<F>(p, ev[(size_t)spec.fac[term][i] * kThreads]);
                        s = fe_add<F>(s, p);
                    }
                    acc[t] = fe_add<F>(acc[t], s);
                }
            }
            if (t < D) {  // e_k(t+1) = e_k(t) + (hi_k - lo_k)
#pragma unroll 1
                for (int k = 0; k < nt; k++) ev[(size_t)k * kThreads] = fe_add<F>(ev[(size_t)k * kThreads], dv[(size_t)k * kThreads]);
            }
        }
    }
    if (WIDE) {
#pragma unroll 1
        for (int t = 0; t <= D; t++) acc[t] = accw_reduce<F>(accw_at(accw, t));
        __syncthreads();
    }
    reduce_publish<F, D + 1>(acc, ra);
}

}  // namespace
}  // namespace zk
