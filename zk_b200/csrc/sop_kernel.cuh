// sop_kernel.cuh — the sum-of-products round kernel (see kernels_sop.cu for what it computes and why).
// Kept in its own header so that tests/cpp/test_sop_kernel_host.cpp can replay the very same source on the host,
// thread by thread, with host_field.hpp standing in for the device arithmetic: the index conventions (pairs, the
// in-place quadruple fold, term bookkeeping) are then checked on the CPU, without a GPU.
// Needs in scope: Fe / FixedMul / TablePtrs / SopSpec (kernels.h), FixedMulF64Sel, the fe_* and ld/st functions
// (field.cuh, field_f64.cuh on the device), kThreads, ReduceArgs and reduce_publish (reduce.cuh on the device), the
// Accw accumulators (accw.cuh on the device), and the two hooks of the dynamic work distribution (DYN only):
// sop_fetch_chunk(ra) = atomicAdd on the launch's work counter, sop_bcast_lane0(v) = warp broadcast from lane 0.
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
// DYN: warps take their chunks of 32 items from a global counter (the WarpChunks of the product kernels) instead of a
// static stride: the schedulers favour some warps and a static split leaves the others to run out the launch alone.
//
// Item layout: e_k = lo_k and d_k = hi_k - lo_k of every table sit in shared memory (2 elements per table and thread,
// the terms index them with run-time table numbers).  spec.n_virt "virtual tables" follow the real ones: table
// n_tables + v is the SUM of tables virt_a[v] and virt_b[v] (e and d add) — how the launcher passes a common factor,
// add.Wb + add.Wc = add.(Wb + Wc): one product less per evaluation point, the same field element by distributivity.
// A term walks its factors once with the running products of ALL D+1 evaluation points in registers
// (e_k(t+1) = e_k(t) + d_k between points): D+1 independent multiplications in flight per factor.
// TOOM (D == 3, every term of at most 3 factors; the launcher checks): the four points are 0, 1, -1 and "infinity" (the
// coefficient of t^3) instead of 0, 1, 2, 3 — e_k(-1) = lo_k - d_k, e_k(inf) = d_k — and a term of fewer than three factors
// has no t^3 coefficient, so it is not multiplied out at infinity at all: the GKR layer add.(Wb + Wc) + mul.Wb.Wc costs 8
// products per item instead of 9 (11 instead of 12 in round 0).  The last block maps the four sums back to S(0..3) with
// exact field arithmetic (reduce.cuh: toom_to_evals, the product kernels' own), so the published values are the same.
template <class F, int D, bool FOLD, bool F64 = false, bool WIDE = false, bool DYN = false, bool TOOM = false>
__global__ void __launch_bounds__(kThreads)
    sop_round_kernel(TablePtrs tabs, const __grid_constant__ SopSpec spec, uint64_t q,
                     const __grid_constant__ FixedMul rtab, const __grid_constant__ FixedMulF64Sel rtab64, ReduceArgs ra) {
    static_assert(!F64 || FOLD, "the FP64 variant only changes the folds");
    static_assert(!TOOM || D == 3, "the Toom point set is wired for cubics");
    const bool skip1 = FOLD && ra.skip1 != 0;
    extern __shared__ __align__(32) uint4 sop_smem[];  // Fe [2 * (n_tables + n_virt)][kThreads]: e_k then d_k; WIDE: + accumulators
    Fe* const ev = reinterpret_cast<Fe*>(sop_smem) + threadIdx.x;
    const int nt = spec.n_tables, nv = spec.n_virt, ntv = nt + nv;
    Fe* const dv = ev + (size_t)ntv * kThreads;
    Accw accw{};
    if (WIDE) {
        uint4* const accw_all = sop_smem + (size_t)2 * ntv * kThreads * (sizeof(Fe) / sizeof(uint4));
        accw_zero(accw_all, D + 1);
        __syncthreads();
        accw = accw_base(accw_all, D + 1);
    }
    Fe acc[D + 1];
#pragma unroll
    for (int t = 0; t <= D; t++) acc[t] = fe_zero<F>();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    constexpr unsigned kWarpsPerBlock = kThreads / 32;
    // chunk = 32 consecutive items; static: chunk c, c + W, c + 2W .. with W = all warps of the grid
    uint32_t c = blockIdx.x * kWarpsPerBlock + warp, cn = c + gridDim.x * kWarpsPerBlock, fetched = 0;
#pragma unroll 1
    while ((uint64_t)c * 32 < q) {
        if constexpr (DYN) {  // ask for the chunk after next (a whole chunk ahead: the atomic's latency is hidden)
            if (lane == 0) fetched = sop_fetch_chunk(ra);
        }
        const uint64_t j = (uint64_t)c * 32 + lane;
        if (j < q) {
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
#pragma unroll 1
            for (int v = 0; v < nv; v++) {
                const int a = spec.virt_a[v], b = spec.virt_b[v];
                ev[(size_t)(nt + v) * kThreads] = fe_add<F>(ev[(size_t)a * kThreads], ev[(size_t)b * kThreads]);
                dv[(size_t)(nt + v) * kThreads] = fe_add<F>(dv[(size_t)a * kThreads], dv[(size_t)b * kThreads]);
            }
#pragma unroll 1
            for (int term = 0; term < spec.n_terms; term++) {
                const int len = (int)spec.len[term];
                // Toom: point index 2 is t = -1, index 3 is "infinity", where only a term of exactly D factors has a value
                const bool inf = !TOOM || len == D;
                Fe p[D + 1];
                {
                    const int f0 = spec.fac[term][0];
                    const Fe d = dv[(size_t)f0 * kThreads];
                    p[0] = ev[(size_t)f0 * kThreads];
                    if (TOOM) {
                        p[1] = fe_add<F>(p[0], d);
                        p[2] = fe_sub<F>(p[0], d);
                        p[3] = d;
                    } else {
#pragma unroll
                        for (int t = 1; t <= D; t++) p[t] = fe_add<F>(p[t - 1], d);
                    }
                }
#pragma unroll 1
                for (int i = 1; i < len; i++) {
                    const int fi = spec.fac[term][i];
                    const Fe e0 = ev[(size_t)fi * kThreads];
                    const Fe d = dv[(size_t)fi * kThreads];
                    const bool last = (i == len - 1);
                    Fe e = e0;
#pragma unroll
                    for (int t = 0; t <= D; t++) {
                        if (TOOM) {
                            if (t == 1) e = fe_add<F>(e0, d);
                            if (t == 2) e = fe_sub<F>(e0, d);
                            if (t == 3) e = d;
                        } else if (t > 0) {
                            e = fe_add<F>(e, d);
                        }
                        if ((t == 1 && skip1) || (TOOM && t == 3 && !inf)) continue;
                        if (WIDE && last) {
                            uint32_t w[16];
                            fe_mul_wide(w, p[t], e);
                            accw_add16(accw_at(accw, t), w);
                        } else {
                            p[t] = fe_mul<F>(p[t], e);
                        }
                    }
                }
                if (WIDE) {
                    if (len == 1) {
#pragma unroll
                        for (int t = 0; t <= D; t++)
                            if (!(t == 1 && skip1) && !(TOOM && t == 3 && !inf)) accw_add_hi(accw_at(accw, t), p[t]);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t <= D; t++)
                        if (!(t == 1 && skip1) && !(TOOM && t == 3 && !inf)) acc[t] = fe_add<F>(acc[t], p[t]);
                }
            }
        }
        if constexpr (DYN) {  // bottom of a chunk (all lanes converged)
            const uint32_t cnn = sop_bcast_lane0(fetched) + 2 * gridDim.x * kWarpsPerBlock;
            c = cn;
            cn = cnn;
        } else {
            c += gridDim.x * kWarpsPerBlock;
        }
    }
    if (WIDE) {
#pragma unroll 1
        for (int t = 0; t <= D; t++) acc[t] = accw_reduce<F>(accw_at(accw, t));
        __syncthreads();
    }
    reduce_publish<F, D + 1, TOOM>(acc, ra);
}

}  // namespace
}  // namespace zk
