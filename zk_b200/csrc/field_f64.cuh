// field_f64.cuh — multiplication by a launch-wide constant on the FP64 pipe (sm_100a: 64 DFMA/clk/SM nominal).
//
// Why: every kernel of the sumcheck path is bound by IMAD.WIDE.U32, which issues at HALF rate on sm_100a
// (30.5/clk/SM measured, tools/ubench.cu) while the FP64 pipe idles.  The fold multiplies every table entry by
// the SAME challenge r (sumcheck/src/prover.rs:64 -> polynomial/src/multilinear/evaluation_form.rs:68), so the
// host can precompute r's multiples and the device only needs EXACT small-integer dot products — which a
// double-precision FMA delivers as long as every partial sum stays below 2^53:
//
//     x        = sum_{i<16} h_i 2^(16 i)                    h_i = 16-bit halves of x's limbs
//     T_i      = r 2^(16 i + 32) mod p = sum_{j<8} T_ij 2^(32 j)       (host, canonical 32-bit limbs as doubles)
//     col_j    = sum_i h_i T_ij            < 16 (2^16-1)(2^32-1) < 2^52   -> 128 exact DFMAs
//     V        = sum_j col_j 2^(32 j)      < 2^276,   V == x r 2^32 (mod p)
//     result   = (V + m p) / 2^32          < 2^244 + p < 2p : ONE Montgomery row (6 wide multiplies), one
//                conditional subtraction -> the canonical residue x r mod p.
//
// For x = aR (Montgomery form) that is (a r)R — bit-identical to ark-ff's `x * r` / fe_mul(x, rR) / fe_mul_fixed.
// The accumulators start at 2^52, so the column sums come out with the integer in the mantissa (no F2I
// conversions); the halves enter through the 2^52 + h bit pattern and one exact subtraction.
// Cost: 144 FP64-pipe instructions + 6 wide multiplies, instead of 76 wide multiplies.
// The arithmetic is plain IEEE-754 binary64 with every intermediate an integer < 2^53, so the same code runs
// bit-identically on the host (tests/cpp/test_f64_fold.cpp checks it against the word-serial host multiplier).
#pragma once
#include <cstdint>
#include <cstring>

#include "field.cuh"

namespace zk {

struct alignas(16) FixedMulF64 {
    double t[16][8];  // t[i][j] = limb j of r * 2^(16 i + 32) mod p
};

// Two identical copies, indexed in kernels by a value that is always 0 but loop-variant in ptxas's eyes: with a
// loop-invariant address ptxas hoists all 128 table entries out of the loop, runs out of uniform registers and
// spills them to local memory (see `ksel` in round_kernel).
struct FixedMulF64Sel {
    FixedMulF64 t[2];
};

namespace detail {
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double f64_from_bits(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ uint32_t f64_lo(double d) { return (uint32_t)__double2loint(d); }
__device__ __forceinline__ uint32_t f64_hi(double d) { return (uint32_t)__double2hiint(d); }
__device__ __forceinline__ double f64_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
inline double f64_from_bits(uint32_t hi, uint32_t lo) {
    uint64_t u = ((uint64_t)hi << 32) | lo;
    double d;
    std::memcpy(&d, &u, 8);
    return d;
}
inline uint32_t f64_lo(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (uint32_t)u; }
inline uint32_t f64_hi(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (uint32_t)(u >> 32); }
inline double f64_fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
#endif
}  // namespace detail

// The two 16-bit halves of a limb as doubles.  Device: one I2F.F64.U16 each (conversion unit, the half is picked
// by the instruction's .H0/.H1 selector).  Host: the 2^52 + h bit pattern and one exact subtraction.  All exact.
#ifdef __CUDACC__
__host__ __device__
#endif
inline void f64_halves(uint32_t x, double& lo, double& hi) {
#if defined(__CUDA_ARCH__)
    asm("{.reg .u16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.rn.f64.u16 %0, l;\n\tcvt.rn.f64.u16 %1, h;}" : "=d"(lo), "=d"(hi) : "r"(x));
#else
    const double two52 = 4503599627370496.0;
    lo = detail::f64_from_bits(0x43300000u, x & 0xffffu) - two52;
    hi = detail::f64_from_bits(0x43300000u, x >> 16) - two52;
#endif
}

// col[k][j] = 2^52 + sum_i h_i(x[k]) * tab.t[i][j]   (exact) for NX operands at once: the operands share every
// table entry (one uniform-register load feeds NX DFMAs).
template <int NX>
#ifdef __CUDACC__
__host__ __device__
#endif
inline void f64_columns_n(double (*col)[8], const uint32_t* const* x, const FixedMulF64& tab,
                          const uint32_t* const* addend = nullptr) {
    const double two52 = 4503599627370496.0;
    // `addend` (optional): a[k][0..6] * 2^32 is added for free by starting column j at 2^52 + a[k][j-1] — that bit
    // pattern is just (0x43300000 : a[k][j-1]).  Still exact: the products leave 2^36 of headroom below 2^52.
#pragma unroll
    for (int k = 0; k < NX; k++)
#pragma unroll
        for (int j = 0; j < 8; j++) col[k][j] = (addend && j > 0) ? detail::f64_from_bits(0x43300000u, addend[k][j - 1]) : two52;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double lo[NX], hi[NX];
#pragma unroll
        for (int k = 0; k < NX; k++) f64_halves(x[k][i], lo[k], hi[k]);
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int k = 0; k < NX; k++) col[k][j] = detail::f64_fma(lo[k], tab.t[2 * i][j], col[k][j]);
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int k = 0; k < NX; k++) col[k][j] = detail::f64_fma(hi[k], tab.t[2 * i + 1][j], col[k][j]);
    }
}
#ifdef __CUDACC__
__host__ __device__
#endif
inline void f64_columns(double* col, const uint32_t* x, const FixedMulF64& tab) {
    const uint32_t* xs[1] = {x};
    f64_columns_n<1>(reinterpret_cast<double(*)[8]>(col), xs, tab);
}

#ifdef __CUDACC__
// column sums (2^52-biased doubles) -> canonical residue
template <class F>
__device__ __forceinline__ Fe f64_columns_reduce(const double* col, uint32_t top_addend = 0, bool twice = false) {
    // the integer part of column j sits in the low 52 bits of the pattern: 32 bits at limb j, 20 bits at limb j+1
    uint32_t V[9];
    V[0] = detail::f64_lo(col[0]);
    asm("add.cc.u32 %0,%8,%16;\n\taddc.cc.u32 %1,%9,%17;\n\taddc.cc.u32 %2,%10,%18;\n\taddc.cc.u32 %3,%11,%19;\n\t"
        "addc.cc.u32 %4,%12,%20;\n\taddc.cc.u32 %5,%13,%21;\n\taddc.cc.u32 %6,%14,%22;\n\taddc.u32 %7,%15,%23;"
        : "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), "=r"(V[8])
        : "r"(detail::f64_hi(col[0]) & 0xfffffu), "r"(detail::f64_hi(col[1]) & 0xfffffu),
          "r"(detail::f64_hi(col[2]) & 0xfffffu), "r"(detail::f64_hi(col[3]) & 0xfffffu),
          "r"(detail::f64_hi(col[4]) & 0xfffffu), "r"(detail::f64_hi(col[5]) & 0xfffffu),
          "r"(detail::f64_hi(col[6]) & 0xfffffu), "r"(detail::f64_hi(col[7]) & 0xfffffu), "r"(detail::f64_lo(col[1])),
          "r"(detail::f64_lo(col[2])), "r"(detail::f64_lo(col[3])), "r"(detail::f64_lo(col[4])),
          "r"(detail::f64_lo(col[5])), "r"(detail::f64_lo(col[6])), "r"(detail::f64_lo(col[7])), "r"(top_addend));
    detail::redc_row_inplace<F, false>(V);  // columns 0..8 -> value in columns 1..8, < 2p (< 3p with an addend)
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = V[1 + i];
    r = fe_reduce_once<F>(r);
    return twice ? fe_reduce_once<F>(r) : r;
}

// x (any value < 2^256) times the constant behind `tab`, fully reduced.
template <class F>
__device__ __forceinline__ Fe fe_mul_fixed_f64(const Fe& x, const FixedMulF64& tab) {
    double col[8];
    f64_columns(col, x.v, tab);
    return f64_columns_reduce<F>(col);
}
// two products by the same constant, table loads shared
template <class F>
__device__ __forceinline__ void fe_mul_fixed_f64_x2(Fe& a, Fe& b, const FixedMulF64& tab) {
    double col[2][8];
    const uint32_t* xs[2] = {a.v, b.v};
    f64_columns_n<2>(col, xs, tab);
    a = f64_columns_reduce<F>(col[0]);
    b = f64_columns_reduce<F>(col[1]);
}

// fold with the FP64 multiples of the challenge: l - r*(l - h)
template <class F>
__device__ __forceinline__ Fe fe_fold_fixed_f64(const Fe& l, const Fe& h, const FixedMulF64& tab) {
    return fe_sub<F>(l, fe_mul_fixed_f64<F>(fe_sub<F>(l, h), tab));
}
// The two folds of one fused-round item: lo = fold(x0, x2), hi = fold(x1, x3), written as l + r (h - l) — the same
// field element as the reference's l - r (l - h) — so that l can ride along as the columns' start value:
// V = l 2^32 + sum_j col_j 2^(32j) < 2^288, (V + m p) / 2^32 < 2^244 + l + p < 3p: two conditional subtractions.
// l is dead before the first DFMA (16 registers less across the dot products than subtracting afterwards).
template <class F>
__device__ __forceinline__ void fe_fold_fixed_f64_x2(Fe& lo, Fe& hi, const Fe& x0, const Fe& x1, const Fe& x2, const Fe& x3,
                                                     const FixedMulF64& tab) {
    const Fe d0 = fe_sub<F>(x2, x0), d1 = fe_sub<F>(x3, x1);
    double col[2][8];
    const uint32_t* xs[2] = {d0.v, d1.v};
    const uint32_t* as[2] = {x0.v, x1.v};
    const uint32_t top0 = x0.v[7], top1 = x1.v[7];
    f64_columns_n<2>(col, xs, tab, as);
    lo = f64_columns_reduce<F>(col[0], top0, true);
    hi = f64_columns_reduce<F>(col[1], top1, true);
}

template <class F>
__device__ __forceinline__ Fe fe_fold_fixed_f64_add(const Fe& l, const Fe& h, const FixedMulF64& tab) {
    const Fe d = fe_sub<F>(h, l);
    double col[1][8];
    const uint32_t* xs[1] = {d.v};
    const uint32_t* as[1] = {l.v};
    const uint32_t top = l.v[7];
    f64_columns_n<1>(col, xs, tab, as);
    return f64_columns_reduce<F>(col[0], top, true);
}
#endif  // __CUDACC__

}  // namespace zk
