// kernels_sumcheck.cu — the sumcheck prover hot path on sm_100a.
//
// What it replaces in the reference (paths relative to /root/reference):
//   sumcheck/src/prover.rs:49-56   for t in 0..=D: poly.partial_evaluate(0,[t]).prod_reduce().iter().sum()
//   sumcheck/src/prover.rs:64      poly = poly.partial_evaluate(0,[challenge])
//   polynomial/src/multilinear/evaluation_form.rs:40-80 (pair (j, j+N/2), left - a*(left-right))
//   polynomial/src/product_poly.rs:66-74 (element-wise product)
// The reference materialises (D+2)*m table clones, D+1 product vectors and D+1 sum passes per round;
// here one pass per round streams every factor table exactly once:
//   round 0            : round_kernel<FOLD=false> reads N            (sum only)
//   round i >= 1       : round_kernel<FOLD=true>  reads N_{i-1}, writes N_{i-1}/2   (fold at r_{i-1} fused
//                        with the sums of round i: thread j owns T[j], T[j+q], T[j+2q], T[j+3q], q = N_{i-1}/4,
//                        writes the two folded values back to T[j], T[j+q] — in place and race-free)
//   after the last rnd : fold_kernel            2 -> 1
// Integer modular sums are order independent, so any reduction tree is bit-exact with the reference's
// sequential `.sum()`.  Reduction: registers -> warp shuffle -> shared memory -> per-block partials in
// global memory -> the last block to finish (atomic ticket) folds the partials and publishes the
// D+1 evaluations to device memory and to mapped pinned host memory.
#include "kernels.h"
#include "field_f64.cuh"
#include "fold_imma.cuh"
#include "reduce.cuh"
#include "accw.cuh"
#include "host_field.hpp"

#include <cstdlib>

namespace zk {
namespace {

// tuning knobs (resident blocks per SM the register allocator must allow, per degree)
#ifndef ZK_RK_MINBLOCKS_D1
#define ZK_RK_MINBLOCKS_D1 6
#endif
#ifndef ZK_RK_MINBLOCKS_D2
#define ZK_RK_MINBLOCKS_D2 5
#endif
#ifndef ZK_RK_MINBLOCKS_D3
#define ZK_RK_MINBLOCKS_D3 4
#endif
#ifndef ZK_RK_MINBLOCKS_IMMA
#define ZK_RK_MINBLOCKS_IMMA 4  // 5 fits (96 registers without the prefetch) and measured slower: 2.50 against 2.36 ms
#endif

// One hypercube item, factor k of m: lo / hi are the pair values of this factor for the item (after the
// optional fold).  Term t of the round polynomial is prod_k e_k(t), e_k(0) = lo_k, e_k(1) = hi_k,
// e_k(t+1) = e_k(t) + (hi_k - lo_k)  (prover.rs:49-56 evaluates at t = 0..D by a full partial_evaluate +
// prod_reduce each; the values are the same field elements).  Factors are visited sequentially so that only
// the D+1 running products and one factor's pair are live, and `m` is a run-time value: one instantiation per
// degree serves every factor count.
// The LAST multiplication of every term is not reduced: the 512-bit product is added to a 17-word
// per-thread accumulator in shared memory and Montgomery-reduced once per thread at the end
// (sum of products then one REDC == sum of REDCs, exactly, mod p): 64 instead of 112 wide multiplies.
// TOOM (only D == 3 with exactly three factors): the terms are carried at the points (0, 1, -1, inf)
// instead of (0, 1, 2, 3).  After the second factor the running product is a quadratic, fixed by three
// values, so its value at -1 is 2(A(0) + A(inf)) - A(1): one multiplication less per item (7 instead of 8).
// The four sums are mapped back to S(0..3) by exact field arithmetic in the last block (toom_to_evals).
// NO1 (with skip1 always on): the accumulator of t = 1 does not exist; the later points move one slot down.
template <class F, int D, bool TOOM, bool NO1 = false>
__device__ __forceinline__ void item_terms(int k, bool last, bool skip1, Fe lo, Fe hi, Fe* pr, const Accw& accw) {
    if (TOOM) {  // pr[0..3] live at t = 0, 1, -1, inf ; m == 3
        const Fe d = fe_sub<F>(hi, lo);
        if (k == 0) {
            pr[0] = lo; pr[1] = hi; pr[3] = d; pr[2] = fe_sub<F>(lo, d);
        } else if (!last) {
            pr[0] = fe_mul<F>(lo, pr[0]);
            pr[1] = fe_mul<F>(hi, pr[1]);
            pr[3] = fe_mul<F>(d, pr[3]);
            const Fe s02 = fe_add<F>(pr[0], pr[3]);
            pr[2] = fe_sub<F>(fe_add<F>(s02, s02), pr[1]);
        } else {
            uint32_t w[16];
            fe_mul_wide(w, lo, pr[0]); accw_add16(accw, w);
            if (!skip1) { fe_mul_wide(w, hi, pr[1]); accw_add16(accw_at(accw, 1), w); }
            fe_mul_wide(w, fe_sub<F>(lo, d), pr[2]); accw_add16(accw_at(accw, NO1 ? 1 : 2), w);
            fe_mul_wide(w, d, pr[3]); accw_add16(accw_at(accw, NO1 ? 2 : 3), w);
        }
    } else if (k == 0) {
        pr[0] = lo;
        if (D >= 1) pr[1] = hi;
        if (D >= 2) {
            Fe d = fe_sub<F>(hi, lo);
#pragma unroll
            for (int t = 2; t <= D; t++) {
                hi = fe_add<F>(hi, d);
                pr[t] = hi;
            }
        }
        if (last) {  // m == 1: the terms are the table values themselves
#pragma unroll
            for (int t = 0; t <= D; t++)
                if (!(skip1 && t == 1)) accw_add_hi(accw_at(accw, t), pr[t]);
        }
    } else {
        uint32_t w[16];
        if (last) { fe_mul_wide(w, lo, pr[0]); accw_add16(accw, w); } else pr[0] = fe_mul<F>(lo, pr[0]);
        if (D >= 2) lo = fe_sub<F>(hi, lo);  // lo := d
        if (D >= 1) {
            if (last) {
                if (!skip1) { fe_mul_wide(w, hi, pr[1]); accw_add16(accw_at(accw, 1), w); }
            } else {
                pr[1] = fe_mul<F>(hi, pr[1]);
            }
        }
#pragma unroll
        for (int t = 2; t <= D; t++) {
            hi = fe_add<F>(hi, lo);
            if (last) { fe_mul_wide(w, hi, pr[t]); accw_add16(accw_at(accw, t), w); } else pr[t] = fe_mul<F>(hi, pr[t]);
        }
    }
}

// the two folds of one fused-round item: lo = fold(x0, x2), hi = fold(x1, x3) at the launch-wide challenge
template <class F, bool F64>
__device__ __forceinline__ void fold_pair(Fe& lo, Fe& hi, const Fe& x0, const Fe& x1, const Fe& x2, const Fe& x3,
                                          const FixedMul& rtab, const FixedMulF64& rtab64) {
    if (F64) {
        fe_fold_fixed_f64_x2<F>(lo, hi, x0, x1, x2, x3, rtab64);
    } else {
        lo = fe_fold_fixed<F>(x0, x2, rtab);  // the challenge enters as precomputed multiples (fe_mul_fixed)
        hi = fe_fold_fixed<F>(x1, x3, rtab);
    }
}

// ---- work distribution ------------------------------------------------------------------------------------------
// The unit of work is a chunk of 32 consecutive items (one warp iteration).  A static grid-stride split leaves
// ~20 % of the warp slots empty on average (ncu: 3.1-3.2 of 4 warps active per scheduler, uniformly on every
// SM): the schedulers favour some warps, those finish their share early and the rest run out the kernel at low
// occupancy.  So each warp takes its first two chunks statically and every later one from a global counter
// (one atomicAdd per 32 items, requested a whole chunk ahead so its latency is hidden); the last block resets
// the counter together with the ticket.  DYN = false keeps the static split: single-table sums are HBM bound and
// so light per chunk that the counter itself (about 0.8e9 same-address atomics per second, measured) would bind.
template <bool DYN>
struct WarpChunks {  // 32-bit chunk ids: tables of up to 2^37 items
    uint32_t c;        // current chunk
    uint32_t cn_dyn;   // DYN: next chunk (static: c + number of warps, not stored)
    uint32_t fetched;  // DYN
    __device__ __forceinline__ WarpChunks(int warp) {
        c = blockIdx.x * kWarps + warp;
        if (DYN) cn_dyn = c + gridDim.x * kWarps;
        fetched = 0;
    }
    __device__ __forceinline__ uint32_t cn() const { return DYN ? cn_dyn : c + gridDim.x * kWarps; }
    __device__ __forceinline__ static bool live(uint32_t chunk, uint64_t q) { return (uint64_t)chunk * 32 < q; }
    __device__ __forceinline__ void request(const ReduceArgs& ra, int lane) {  // top of a chunk: ask for the chunk after next
        // inline PTX: nvcc turns a plain atomicAdd on a uniform address into a warp-aggregated one (ballot, leader, SHFL of
        // the result), and that shuffle waits for the atomic right here instead of a chunk later (ncu: 2.3 % of the samples)
        if (DYN && lane == 0)
            asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(fetched) : "l"(ra.ticket + kWorkCounterOffset) : "memory");
    }
    __device__ __forceinline__ void advance() {  // bottom of a chunk (all lanes converged)
        if (DYN) {
            const uint32_t cnn = __shfl_sync(0xffffffffu, fetched, 0) + 2 * gridDim.x * kWarps;
            c = cn_dyn;
            cn_dyn = cnn;
        } else {
            c += gridDim.x * kWarps;
        }
    }
};

// ---- the round kernel -----------------------------------------------------------------------------------------
// FP (only with FOLD) = the pipe that folds.  0: the integer pipe (fe_mul_fixed's 76 wide multiplies).  1: the FP64 pipe
// (field_f64.cuh: exact DFMA dot products against 16 host-made multiples of the challenge, one Montgomery row).
// 2 / 3: the INT8 tensor path (fold_imma.cuh: 8 IMMA per warp and fold on the bytes of h - l against a 32 x 32 byte table of
// the challenge's multiples, one Montgomery row; 3 stages the two folds of an item together).  The table travels in the
// bytes of rtab64.t[0]; the per-warp staging area follows the wide accumulators in shared memory.  q must be a multiple
// of 32 (every lane of a warp owns an item: mma.sync).
template <class F, int D, bool FOLD, bool TOOM = false, int FP = 0, bool DYN = false>
__global__ void __launch_bounds__(kThreads, (D <= 1) ? (FOLD ? ZK_RK_MINBLOCKS_D1 : ZK_RK_MINBLOCKS_D1 - 1) : (D == 2 ? ZK_RK_MINBLOCKS_D2 : (FP >= 2 ? ZK_RK_MINBLOCKS_IMMA : ZK_RK_MINBLOCKS_D3)))
    round_kernel(TablePtrs tabs, int m, uint64_t q, uint64_t hoff, const __grid_constant__ FixedMul rtab,
                 const __grid_constant__ FixedMulF64Sel rtab64, ReduceArgs ra) {
    static_assert(!TOOM || D == 3, "the Toom point set is wired for cubics");
    static_assert(FP == 0 || FOLD, "the fold-pipe variants only change the folds");
    constexpr bool F64 = FP == 1, IMMA = FP >= 2;
    constexpr int NF = FP == 3 ? 2 : 1;
    // the tensor-path variants run with the claim known (skip1): no accumulator for t = 1 (8.7 KB of shared memory less)
    constexpr int NPA = IMMA ? D : D + 1;
    extern __shared__ uint4 accw_all[];  // accw_bytes(NPA) [+ kWarps * NF * kImmaStageBytes]
    __shared__ Fe* s_tab[kMaxFactors];
    if (threadIdx.x < kMaxFactors) s_tab[threadIdx.x] = tabs.t[threadIdx.x];
    accw_zero(accw_all, NPA);
    __syncthreads();
    const Accw accw = accw_base(accw_all, NPA);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ImmaTab itab{};
    unsigned char* stage = nullptr;
    if constexpr (IMMA) {
        static_assert(sizeof(FixedMulI8) <= sizeof(FixedMulF64), "the byte table travels in the FP64 table's slot");
        itab = imma_tab_load(reinterpret_cast<const FixedMulI8&>(rtab64.t[0]), lane);
        stage = reinterpret_cast<unsigned char*>(accw_all) + accw_bytes(NPA) + (size_t)warp * NF * kImmaStageBytes;
    }
    auto fold2 = [&](Fe& lo, Fe& hi, const Fe& x0, const Fe& x1, const Fe& x2, const Fe& x3, int ksel) {
        if constexpr (IMMA) {
            if constexpr (NF == 2) {
                Fe o[2];
                const Fe l2[2] = {x0, x1}, h2[2] = {x2, x3};
                fe_fold_imma_n<F, 2>(o, l2, h2, itab, stage, lane);
                lo = o[0];
                hi = o[1];
            } else {
                lo = fe_fold_imma<F>(x0, x2, itab, stage, lane);
                hi = fe_fold_imma<F>(x1, x3, itab, stage, lane);
            }
        } else {
            fold_pair<F, F64>(lo, hi, x0, x1, x2, x3, rtab, rtab64.t[ksel]);
        }
    };
    WarpChunks<DYN> wc(warp);
    // Software pipelining of the global loads.  Round 0 (no fold) keeps the pair of the next (item, factor) in
    // flight while this one is multiplied.  The fused kernel issues the four loads of the next (item, factor)
    // right AFTER this factor's folds (when x0..x3 are dead) so they land during the product multiplications
    // (-4 % at D = 3; at D <= 2 the register budget of the higher-occupancy variants makes it a loss).
    // (a tensor-path build with five resident blocks has no registers for the prefetched quadruple)
    constexpr bool kFoldPrefetch = FOLD && D >= 3 && (!IMMA || ZK_RK_MINBLOCKS_IMMA <= 4);
    Fe n0, n1, n2, n3;
    if ((uint64_t)wc.c * 32 + lane < q) {
        const uint64_t j0 = (uint64_t)wc.c * 32 + lane;
        Fe* T = s_tab[0];
        if (FOLD) {
            if (kFoldPrefetch) {
                n0 = ld_fe_stream(T + j0); n2 = ld_fe_stream(T + j0 + 2 * q); n1 = ld_fe_stream(T + j0 + q); n3 = ld_fe_stream(T + j0 + 3 * q);
            }
        } else {
            n0 = ld_fe_stream(T + j0);
            n1 = ld_fe_stream(T + j0 + hoff);
        }
    }
#pragma unroll 1
    while (WarpChunks<DYN>::live(wc.c, q)) {
        wc.request(ra, lane);
        const uint64_t j = (uint64_t)wc.c * 32 + lane, jn = (uint64_t)wc.cn() * 32 + lane;
        const bool nvalid = jn < q;
        if constexpr (IMMA && TOOM) {
            // Three factors, unrolled, every lane owns an item.  The first factor has no products to hide the next
            // quadruple behind (ncu: 5.8 % of the samples waited for it at the head of the second fold), so table 1's
            // quadruple is requested BEFORE the first fold, into registers the running products do not need yet; table 2's
            // goes out before the second fold, the next item's first quadruple behind the last products as before.
            Fe pr[D + 1], lo, hi;
            Fe* const T0 = s_tab[0];
            Fe* const T1 = s_tab[1];
            Fe* const T2 = s_tab[2];
            const Fe m0 = ld_fe_stream(T1 + j), m2 = ld_fe_stream(T1 + j + 2 * q), m1 = ld_fe_stream(T1 + j + q), m3 = ld_fe_stream(T1 + j + 3 * q);
            // (prefetch.global.L2 of the next item's twelve rows a whole item ahead measured slower: 2.24 against 2.18 ms)
            fold2(lo, hi, n0, n1, n2, n3, 0);
            st_fe(T0 + j, lo);
            st_fe(T0 + j + q, hi);
            n0 = ld_fe_stream(T2 + j); n2 = ld_fe_stream(T2 + j + 2 * q); n1 = ld_fe_stream(T2 + j + q); n3 = ld_fe_stream(T2 + j + 3 * q);
            item_terms<F, D, TOOM, IMMA>(0, false, true, lo, hi, pr, accw);
            fold2(lo, hi, m0, m1, m2, m3, 0);
            st_fe(T1 + j, lo);
            st_fe(T1 + j + q, hi);
            item_terms<F, D, TOOM, IMMA>(1, false, true, lo, hi, pr, accw);
            fold2(lo, hi, n0, n1, n2, n3, 0);
            st_fe(T2 + j, lo);
            st_fe(T2 + j + q, hi);
            if (nvalid) {
                n0 = ld_fe_stream(T0 + jn); n2 = ld_fe_stream(T0 + jn + 2 * q); n1 = ld_fe_stream(T0 + jn + q); n3 = ld_fe_stream(T0 + jn + 3 * q);
            }
            item_terms<F, D, TOOM, IMMA>(2, true, true, lo, hi, pr, accw);
        } else if (j < q) {  // false only in the single ragged chunk of a table with fewer than 32 items
            Fe pr[D + 1];
#pragma unroll 1
            for (int k = 0; k < m; k++) {
                Fe* T = s_tab[k];
                Fe lo, hi;
                // Always 0 — but a loop-variant index in ptxas's eyes: with a loop-invariant address it hoists all 128
                // table entries out of the loop, runs out of uniform registers and spills them to local memory.
                const int ksel = F64 ? (k >> 16) : 0;
                const bool more_k = (k + 1 < m);
                const uint64_t nj = more_k ? j : jn;
                const bool nok = more_k || nvalid;
                Fe* NT = s_tab[more_k ? k + 1 : 0];
                if (FOLD) {  // T has 4q entries: fold (j, j+2q) and (j+q, j+3q) at r, write back to j and j+q
                    if (kFoldPrefetch) {
                        fold2(lo, hi, n0, n1, n2, n3, ksel);
                        st_fe(T + j, lo);
                        st_fe(T + j + q, hi);
                        if (nok) {
                            n0 = ld_fe_stream(NT + nj); n2 = ld_fe_stream(NT + nj + 2 * q);
                            n1 = ld_fe_stream(NT + nj + q); n3 = ld_fe_stream(NT + nj + 3 * q);
                        }
                    } else {
                        Fe x0 = ld_fe_stream(T + j), x2 = ld_fe_stream(T + j + 2 * q);
                        Fe x1 = ld_fe_stream(T + j + q), x3 = ld_fe_stream(T + j + 3 * q);
                        fold2(lo, hi, x0, x1, x2, x3, ksel);
                        st_fe(T + j, lo);
                        st_fe(T + j + q, hi);
                    }
                } else {  // T has 2q entries: the pair is (j, j+q)
                    lo = n0;
                    hi = n1;
                    if (nok) {
                        n0 = ld_fe_stream(NT + nj);
                        n1 = ld_fe_stream(NT + nj + hoff);
                    }
                }
                item_terms<F, D, TOOM, IMMA>(k, k == m - 1, IMMA || (FOLD && ra.skip1 != 0), lo, hi, pr, accw);
            }
        }
        wc.advance();
    }
    Fe acc[D + 1];
#pragma unroll 1
    for (int t = 0; t <= D; t++) {
        if (IMMA && t == 1) acc[t] = fe_zero<F>();  // S(1) comes from the claim (reduce_publish)
        else acc[t] = accw_reduce<F>(accw_at(accw, IMMA && t > 1 ? t - 1 : t));
    }
    __syncthreads();
    reduce_publish<F, D + 1, TOOM>(acc, ra);
}

// grid.y = factor index
template <class F>
__global__ void __launch_bounds__(256) fold_kernel(TablePtrs tabs, uint64_t half, const __grid_constant__ FixedMul rtab) {
    Fe* T = tabs.t[blockIdx.y];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride) {
        Fe l = ld_fe_stream(T + j), h = ld_fe_stream(T + j + half);
        st_fe(T + j, fe_fold_fixed<F>(l, h, rtab));
    }
}

// Generic (any m <= kMaxFactors): one evaluation point per launch,
// sum_j prod_k [lo_k - t (lo_k - hi_k)]   (t given in Montgomery form; is_zero/is_one shortcuts are
// the same field function, evaluation_form.rs:61-62).
template <class F>
__global__ void __launch_bounds__(kThreads)
    eval_at_kernel(TablePtrs tabs, int m, uint64_t half, const __grid_constant__ FixedMul ttab, ReduceArgs ra) {
    Fe acc[1];
    acc[0] = fe_zero<F>();
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
#pragma unroll 1
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < half; j += stride) {
        Fe pr = fe_fold_fixed<F>(ld_fe_stream(tabs.t[0] + j), ld_fe_stream(tabs.t[0] + j + half), ttab);
#pragma unroll 1
        for (int k = 1; k < m; k++)
            pr = fe_mul<F>(pr, fe_fold_fixed<F>(ld_fe_stream(tabs.t[k] + j), ld_fe_stream(tabs.t[k] + j + half), ttab));
        acc[0] = fe_add<F>(acc[0], pr);
    }
    reduce_publish<F, 1>(acc, ra);
}

template <class F>
__global__ void __launch_bounds__(kThreads) product_sum_kernel(TablePtrs tabs, int m, uint64_t n, ReduceArgs ra) {
    Fe acc[1];
    acc[0] = fe_zero<F>();
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
#pragma unroll 1
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
        Fe pr = ld_fe_stream(tabs.t[0] + j);
#pragma unroll 1
        for (int k = 1; k < m; k++) pr = fe_mul<F>(pr, ld_fe_stream(tabs.t[k] + j));
        acc[0] = fe_add<F>(acc[0], pr);
    }
    reduce_publish<F, 1>(acc, ra);
}

// ZK_B200_SCHED=static keeps the grid-stride split (A/B measurements); default: dynamic chunks.
inline bool dynamic_chunks() {
    static const bool on = [] {
        const char* e = std::getenv("ZK_B200_SCHED");
        return !(e && e[0] == 's');
    }();
    return on;
}
template <class F>
Fe small_constant(unsigned t);  // Montgomery form of small integer t (host side)


// Which pipe folds: FP64 (field_f64.cuh) for items of two or more factors, where the product multiplications keep
// the integer-multiply pipe busy (measured -3 % on the fused degree-3 step, -1.5 % at degree 2); the integer pipe
// for a single table (HBM bound: the longer FP64 instruction stream only costs).  ZK_B200_FOLD_PIPE=int|f64 forces one.
inline bool fold_on_f64(int m) {
    static const int forced = [] {
        const char* e = std::getenv("ZK_B200_FOLD_PIPE");
        return !e ? -1 : (e[0] == 'i' && e[1] == 'n' ? 0 : 1);
    }();
    return forced >= 0 ? forced == 1 : m >= 2;
}
// The INT8 tensor path (fold_imma.cuh) for the degree-3, three-factor fused step with the claim known (the prover's
// rounds): measured 2.18 ms against 2.44 ms (FP64 folds) for the first fused step of the 2^26 proof, bit-exact.
// ZK_B200_FOLD_PIPE=imma2 (default: the two folds of an item staged together) | imma (one at a time: 2.52 ms) |
// f64 | int (the other pipes).  Returns the kernel's FP parameter, 0 = off.
inline int fold_on_imma() {
    static const int mode = [] {
        const char* e = std::getenv("ZK_B200_FOLD_PIPE");
        if (!e) return 3;
        if (e[0] != 'i' || e[1] != 'm') return 0;
        return e[4] == '2' ? 3 : 2;
    }();
    return mode;
}
// host: the byte table of the challenge's multiples in IMMA fragment order, carried in the first FP64 table's bytes
template <class F>
FixedMulF64Sel make_fixed_i8(const Fe& r) {
    static_assert(sizeof(FixedMulI8) <= sizeof(FixedMulF64), "the byte table travels in the FP64 table's slot");
    host::Field HF(F::ID);
    host::El rm;
    std::memcpy(rm.v, r.v, 32);
    FixedMulI8 t8;
    fixed_mul_table_i8(HF, rm, &t8);
    FixedMulF64Sel t{};
    std::memcpy(&t.t[0], &t8, sizeof(t8));
    return t;
}

template <class F, int D, bool FOLD, bool TOOM, int FP, bool DYN>
cudaError_t do_round_v(const TablePtrs& tabs, int m, uint64_t q, const Fe& r, const ReduceScratch& s, cudaStream_t st, const Fe* claim, uint64_t hoff) {
    constexpr bool F64 = FP == 1, IMMA = FP >= 2;
    const FixedMul tab = (FOLD && FP == 0) ? make_fixed<F>(r) : FixedMul{};
    const FixedMulF64Sel tab64 = F64 ? make_fixed_f64<F>(r) : (IMMA ? make_fixed_i8<F>(r) : FixedMulF64Sel{});
    constexpr size_t smem = IMMA ? accw_bytes(D) + (size_t)kWarps * (FP == 3 ? 2 : 1) * kImmaStageBytes : accw_bytes(D + 1);
    static PerDeviceCache cache;
    const int bpsm = per_device(cache, [] {
        cudaError_t e = cudaFuncSetAttribute(round_kernel<F, D, FOLD, TOOM, FP, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, round_kernel<F, D, FOLD, TOOM, FP, DYN>, kThreads, smem) != cudaSuccess || nb < 1) nb = 1;
        return nb;
    });
    if (bpsm <= 0) return (cudaError_t)(-bpsm);
    unsigned grid = grid_for(q, kThreads, s.num_sms, bpsm);
    ReduceArgs ra = make_ra(s, 0);
    if (FOLD && claim && D >= 1) {
        ra.skip1 = 1;
        ra.claim = *claim;
    }
    round_kernel<F, D, FOLD, TOOM, FP, DYN><<<grid, kThreads, smem, st>>>(tabs, m, q, hoff ? hoff : q, tab, tab64, ra);
    return cudaGetLastError();
}
template <class F, int D, bool FOLD, bool TOOM = false>
cudaError_t do_round(const TablePtrs& tabs, int m, uint64_t q, const Fe& r, const ReduceScratch& s, cudaStream_t st, const Fe* claim, uint64_t hoff) {
    const bool dyn = m >= 2 && dynamic_chunks();
    if constexpr (FOLD && D == 3 && TOOM) {
        const int imma = fold_on_imma();
        if (imma && dyn && q % 32 == 0 && claim)
            return imma == 3 ? do_round_v<F, D, FOLD, TOOM, 3, true>(tabs, m, q, r, s, st, claim, hoff) : do_round_v<F, D, FOLD, TOOM, 2, true>(tabs, m, q, r, s, st, claim, hoff);
    }
    if (FOLD && fold_on_f64(m))
        return dyn ? do_round_v<F, D, FOLD, TOOM, FOLD, true>(tabs, m, q, r, s, st, claim, hoff) : do_round_v<F, D, FOLD, TOOM, FOLD, false>(tabs, m, q, r, s, st, claim, hoff);
    return dyn ? do_round_v<F, D, FOLD, TOOM, false, true>(tabs, m, q, r, s, st, claim, hoff) : do_round_v<F, D, FOLD, TOOM, false, false>(tabs, m, q, r, s, st, claim, hoff);
}
template <class F, bool FOLD>
cudaError_t do_round_deg(const TablePtrs& tabs, int m, int degree, uint64_t q, const Fe& r, const ReduceScratch& s,
                         cudaStream_t st, const Fe* claim = nullptr, uint64_t hoff = 0) {
    switch (degree) {
        case 1: return do_round<F, 1, FOLD>(tabs, m, q, r, s, st, claim, hoff);
        case 2: return do_round<F, 2, FOLD>(tabs, m, q, r, s, st, claim, hoff);
        case 3: return m == 3 ? do_round<F, 3, FOLD, true>(tabs, m, q, r, s, st, claim, hoff) : do_round<F, 3, FOLD>(tabs, m, q, r, s, st, claim, hoff);
        case 4: return do_round<F, 4, FOLD>(tabs, m, q, r, s, st, claim, hoff);
        default: return cudaErrorInvalidValue;
    }
}

// ---- the latency kernel of the small rounds --------------------------------------------------------------------
// From about 2^17 items down a round no longer fills the GPU: with one item per thread its duration is the latency of ONE
// item (six folds and seven products back to back, about 3400 dependent-ish instructions = 9 us) plus the wide-accumulator
// and grid reductions, 23 - 27 us per launch whatever the size — eighteen such rounds are 6 % of a 2^26 proof.  Here EIGHT
// lanes share an item: lanes 0 .. 2m-1 each fold one pair of one factor (lane 2k: (j, j+2q) -> lo_k, lane 2k+1:
// (j+q, j+3q) -> hi_k, written back in place like the streaming kernel), the folded values travel by warp shuffle, and lane
// t <= D multiplies its own evaluation point e_k(t) = lo_k + t (hi_k - lo_k) through the factors: the critical path is one
// fold and m-1 products.  Plain reduced products (no wide accumulators to drain), a shuffle + shared-memory reduction of one
// element per lane, the last block publishes exactly like reduce_publish.  Same field elements as the streaming kernel
// (any exact evaluation of prover.rs:49-56 / :64 is), so which kernel a round takes is invisible in the proof.
// m <= 4 (2m fold lanes) and D <= 4; other shapes stay on the streaming kernel.
constexpr int kSmallGroup = 8;
// The polynomial is given as a sum of products (SopSpec; a ProductPoly is the single term 0.1...m-1), so the same kernel
// serves the small rounds of the sum-of-products prover (kernels_sop.cu): lane t walks the terms, fetching lo_f / hi_f of
// each factor from the lanes that folded them.
template <class F, int D>
__global__ void __launch_bounds__(kThreads)
    round_small_kernel(TablePtrs tabs, const __grid_constant__ SopSpec spec, uint64_t q, Fe r_in, ReduceArgs ra) {
    const int m = spec.n_tables;
    constexpr int NP = D + 1;
    __shared__ Fe sh[kWarps][kSmallGroup];
    __shared__ Fe sh_fin[kThreads / kSmallGroup][kSmallGroup];
    __shared__ unsigned s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane & (kSmallGroup - 1), base = lane & ~(kSmallGroup - 1);
    Fe r = r_in;  // a multiplier read from the constant bank makes ptxas split the wide multiplies: keep it in registers
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("" : "+r"(r.v[i]));
    const bool skip1 = ra.skip1 != 0;
    const int k_mine = g >> 1, h_mine = g & 1;
    const uint64_t items_per_grid = (uint64_t)gridDim.x * (kThreads / kSmallGroup);
    Fe acc = fe_zero<F>();
    // all lanes of a warp run the same number of iterations (the shuffles are warp-wide); `live` masks the ragged end
#pragma unroll 1
    for (uint64_t j0 = (uint64_t)blockIdx.x * (kThreads / kSmallGroup) + warp * (32 / kSmallGroup); j0 < q; j0 += items_per_grid) {
        const uint64_t j = j0 + (lane >> 3);
        const bool live = j < q;
        Fe v = fe_zero<F>();
        if (live && g < 2 * m) {
            Fe* T = tabs.t[k_mine] + j + (h_mine ? q : 0);
            const Fe a = ld_fe_stream(T), b = ld_fe_stream(T + 2 * q);
            v = fe_fold<F>(a, b, r);  // a - r (a - b): evaluation_form.rs:68
            st_fe(T, v);
        }
        Fe tot = fe_zero<F>();
#pragma unroll 1
        for (int term = 0; term < spec.n_terms; term++) {
            Fe pr = fe_zero<F>();
#pragma unroll 1
            for (int fi = 0; fi < (int)spec.len[term]; fi++) {
                const int k = spec.fac[term][fi];
                Fe lo, hi;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    lo.v[i] = __shfl_sync(0xffffffffu, v.v[i], base + 2 * k);
                    hi.v[i] = __shfl_sync(0xffffffffu, v.v[i], base + 2 * k + 1);
                }
                // e = e_k(g): lo, hi, hi + d, hi + 2d, ...
                Fe e = (g == 0) ? lo : hi;
                if (D >= 2) {
                    const Fe d = fe_sub<F>(hi, lo);
#pragma unroll
                    for (int t = 2; t <= D; t++) {
                        const Fe e2 = fe_add<F>(e, d);
                        if (g >= t) e = e2;
                    }
                }
                pr = (fi == 0) ? e : fe_mul<F>(pr, e);
            }
            tot = (term == 0) ? pr : fe_add<F>(tot, pr);
        }
        if (live && g < NP && !(skip1 && g == 1)) acc = fe_add<F>(acc, tot);
    }
    // warp: the four groups' lanes of the same point
#pragma unroll
    for (int off = kSmallGroup; off < 32; off <<= 1) {
        Fe o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.v[i] = __shfl_xor_sync(0xffffffffu, acc.v[i], off);
        acc = fe_add<F>(acc, o);
    }
    if (lane < kSmallGroup) sh[warp][lane] = acc;
    __syncthreads();
    if (gridDim.x == 1) {  // the last rounds of every proof: one block, nothing to hand over through global memory
        if (threadIdx.x < kSmallGroup) {
            Fe v = sh[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < kWarps; w++) v = fe_add<F>(v, sh[w][threadIdx.x]);
            sh_fin[0][threadIdx.x] = v;
        }
        __syncthreads();
        finish_last_block<F, NP, false>(&sh_fin[0][0], ra);
        return;
    }
    if (threadIdx.x < NP) {
        Fe v = sh[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kWarps; w++) v = fe_add<F>(v, sh[w][threadIdx.x]);
        st_fe(ra.block_partials + (size_t)blockIdx.x * NP + threadIdx.x, v);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned tk = atomicAdd(ra.ticket, 1u);
        s_last = (tk == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {   // last block: thread (row, t) sums the partials of blocks row, row + 16, ... for point t
        const int t = threadIdx.x & (kSmallGroup - 1), row = threadIdx.x >> 3;
        Fe v = fe_zero<F>();
        if (t < NP)
            for (unsigned b = row; b < gridDim.x; b += kThreads / kSmallGroup) v = fe_add<F>(v, ld_fe_cg(ra.block_partials + (size_t)b * NP + t));
        sh_fin[row][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < kSmallGroup) {
        Fe v = sh_fin[0][threadIdx.x];
#pragma unroll 1
        for (int row = 1; row < kThreads / kSmallGroup; row++) v = fe_add<F>(v, sh_fin[row][threadIdx.x]);
        sh[0][threadIdx.x] = v;
    }
    __syncthreads();
    finish_last_block<F, NP, false>(&sh[0][0], ra);
}

// Items at or below which the fused step takes the latency kernel (ZK_B200_SMALL_Q overrides; 0 disables it).
inline uint64_t small_q_threshold() {
    static const uint64_t v = [] {
        const char* e = std::getenv("ZK_B200_SMALL_Q");
        return e ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)1 << 13;  // measured crossover on B200: 2^13 items (profiles/r02_small_round_kernel_ab.txt)
    }();
    return v;
}
template <class F, int D>
cudaError_t do_round_small_d(const TablePtrs& tabs, const SopSpec& spec, uint64_t q, const Fe& r, const ReduceScratch& s, cudaStream_t st, const Fe* claim) {
    const uint64_t per_block = kThreads / kSmallGroup;
    uint64_t need = (q + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t)s.num_sms * 16 < (uint64_t)kMaxGridBlocks ? (uint64_t)s.num_sms * 16 : (uint64_t)kMaxGridBlocks;
    if (need < 1) need = 1;
    ReduceArgs ra = make_ra(s, 0);
    if (claim) {
        ra.skip1 = 1;
        ra.claim = *claim;
    }
    round_small_kernel<F, D><<<(unsigned)(need < cap ? need : cap), kThreads, 0, st>>>(tabs, spec, q, r, ra);
    return cudaGetLastError();
}
template <class F>
cudaError_t do_round_small(const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t q, const Fe& r, const ReduceScratch& s, cudaStream_t st,
                           const Fe* claim) {
    switch (degree) {
        case 1: return do_round_small_d<F, 1>(tabs, spec, q, r, s, st, claim);
        case 2: return do_round_small_d<F, 2>(tabs, spec, q, r, s, st, claim);
        case 3: return do_round_small_d<F, 3>(tabs, spec, q, r, s, st, claim);
        case 4: return do_round_small_d<F, 4>(tabs, spec, q, r, s, st, claim);
        default: return cudaErrorInvalidValue;
    }
}
inline SopSpec product_as_spec(int m) {  // ProductPoly = the single term 0.1...(m-1)
    SopSpec sp{};
    sp.n_tables = m;
    sp.n_terms = 1;
    sp.len[0] = (uint8_t)m;
    for (int k = 0; k < m; k++) sp.fac[0][k] = (uint8_t)k;
    return sp;
}

template <class F>
cudaError_t round_poly_dispatch(const TablePtrs& tabs, int m, int degree, uint64_t half, const ReduceScratch& s,
                                cudaStream_t st, int* launches) {
    if (has_fused_path(m, degree)) { ++*launches; return do_round_deg<F, false>(tabs, m, degree, half, Fe{}, s, st); }
    const int bpsm = blocks_per_sm(eval_at_kernel<F>, kThreads);
    unsigned grid = grid_for(half, kThreads, s.num_sms, bpsm);
    for (int t = 0; t <= degree; t++) {
        ReduceArgs ra = make_ra(s, t);
        if (t != degree) ra.seq = 0;  // only the last launch of the group publishes the completion flag
        eval_at_kernel<F><<<grid, kThreads, 0, st>>>(tabs, m, half, make_fixed<F>(host_small_mont<F>((unsigned)t)), ra);
        ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
template <class F>
cudaError_t fold_dispatch(const TablePtrs& tabs, int m, uint64_t half, const Fe& r, const ReduceScratch* s,
                          cudaStream_t st, int* launches) {
    int num_sms = s ? s->num_sms : 148;
    uint64_t need = (half + 255) / 256;
    uint64_t cap = (uint64_t)num_sms * 8;
    dim3 grid((unsigned)(need < cap ? (need ? need : 1) : cap), (unsigned)m);
    fold_kernel<F><<<grid, 256, 0, st>>>(tabs, half, make_fixed<F>(r));
    ++*launches;
    return cudaGetLastError();
}
template <class F>
cudaError_t fold_round_poly_dispatch(const TablePtrs& tabs, int m, int degree, uint64_t n_prev, const Fe& r,
                                     const ReduceScratch& s, cudaStream_t st, int* launches, const Fe* claim) {
    const uint64_t q = n_prev / 4;
    if (has_fused_path(m, degree) && 2 * m <= kSmallGroup && q <= small_q_threshold()) {
        ++*launches;
        return do_round_small<F>(tabs, product_as_spec(m), degree, q, r, s, st, claim);
    }
    if (has_fused_path(m, degree)) { ++*launches; return do_round_deg<F, true>(tabs, m, degree, q, r, s, st, claim); }
    cudaError_t e = fold_dispatch<F>(tabs, m, n_prev / 2, r, &s, st, launches);
    if (e != cudaSuccess) return e;
    return round_poly_dispatch<F>(tabs, m, degree, n_prev / 4, s, st, launches);
}
template <class F>
cudaError_t product_sum_dispatch(const TablePtrs& tabs, int m, uint64_t n, const ReduceScratch& s, cudaStream_t st,
                                 int* launches) {
    const int bpsm = blocks_per_sm(product_sum_kernel<F>, kThreads);
    unsigned grid = grid_for(n, kThreads, s.num_sms, bpsm);
    product_sum_kernel<F><<<grid, kThreads, 0, st>>>(tabs, m, n, make_ra(s, 0));
    ++*launches;
    return cudaGetLastError();
}

}  // namespace

bool has_fused_path(int m, int degree) { return m >= 1 && m <= kMaxFactors && degree >= 1 && degree <= 4; }

bool small_round_applies(int n_tables, int degree, uint64_t q) {
    return 2 * n_tables <= kSmallGroup && degree >= 1 && degree <= 4 && q <= small_q_threshold();
}
cudaError_t launch_small_fold_round(int field, const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t q, const Fe& r,
                                    const ReduceScratch& scratch, cudaStream_t stream, int* launches, const Fe* claim) {
    ++*launches;
    return field == Fr381::ID ? do_round_small<Fr381>(tabs, spec, degree, q, r, scratch, stream, claim)
                              : do_round_small<Fr377>(tabs, spec, degree, q, r, scratch, stream, claim);
}

cudaError_t launch_round_poly(int field, const TablePtrs& tabs, int m, int degree, uint64_t half,
                              const ReduceScratch& scratch, cudaStream_t stream, int* launches) {
    return field == Fr381::ID ? round_poly_dispatch<Fr381>(tabs, m, degree, half, scratch, stream, launches)
                              : round_poly_dispatch<Fr377>(tabs, m, degree, half, scratch, stream, launches);
}
cudaError_t launch_round_poly_range(int field, const TablePtrs& tabs, int m, int degree, uint64_t count, uint64_t hoff,
                                    const ReduceScratch& scratch, cudaStream_t stream, int* launches) {
    if (!has_fused_path(m, degree) || count < 1 || hoff < count) return cudaErrorInvalidValue;
    ++*launches;
    return field == Fr381::ID ? do_round_deg<Fr381, false>(tabs, m, degree, count, Fe{}, scratch, stream, nullptr, hoff)
                              : do_round_deg<Fr377, false>(tabs, m, degree, count, Fe{}, scratch, stream, nullptr, hoff);
}
cudaError_t launch_fold(int field, const TablePtrs& tabs, int m, uint64_t half, const Fe& r, cudaStream_t stream,
                        int* launches) {
    return field == Fr381::ID ? fold_dispatch<Fr381>(tabs, m, half, r, nullptr, stream, launches)
                              : fold_dispatch<Fr377>(tabs, m, half, r, nullptr, stream, launches);
}
cudaError_t launch_fold_round_poly(int field, const TablePtrs& tabs, int m, int degree, uint64_t n_prev, const Fe& r,
                                   const ReduceScratch& scratch, cudaStream_t stream, int* launches, const Fe* claim) {
    return field == Fr381::ID ? fold_round_poly_dispatch<Fr381>(tabs, m, degree, n_prev, r, scratch, stream, launches, claim)
                              : fold_round_poly_dispatch<Fr377>(tabs, m, degree, n_prev, r, scratch, stream, launches, claim);
}
cudaError_t launch_product_sum(int field, const TablePtrs& tabs, int m, uint64_t n, const ReduceScratch& scratch,
                               cudaStream_t stream, int* launches) {
    return field == Fr381::ID ? product_sum_dispatch<Fr381>(tabs, m, n, scratch, stream, launches)
                              : product_sum_dispatch<Fr377>(tabs, m, n, scratch, stream, launches);
}

}  // namespace zk
