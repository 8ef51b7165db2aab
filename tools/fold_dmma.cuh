// fold_dmma.cuh — the two folds of a warp's 32 items on the FP64 tensor path (DMMA.8x8x4).  Needs field_f64.cuh.
#pragma once
#include "field_f64.cuh"

namespace zk {
// ---- the same two folds as ONE small dense contraction per warp on the FP64 tensor path (DMMA.8x8x4) -------------------
// The two folds of a warp's 32 items are a [64 x 16] x [16 x 8] matrix product: rows = (fold, item), K = the sixteen
// 16-bit halves of h - l, N = the eight 32-bit limbs of the challenge's multiples T_i (FixedMulF64, unchanged).  As
// mma.sync.m8n8k4.f64 that is 32 DMMA per warp instead of 256 DFMA + 128 uniform loads per thread: an eighth of the issue
// slots for the same arithmetic (exact: every partial sum is an integer below 2^53 whatever the summation order;
// tools/dmma_probe.cu checks the instruction, tools/dmma_fold_probe.cu this function against the integer fold), and the
// tensor pipe works while the schedulers issue other warps' integer multiplications.  Measured stand-alone at the round
// kernel's occupancy (B200): 9.5e10 folds/s against 6.2e10 (DFMA) and 8.6e10 (IMAD.WIDE).
// The operands change layout through a 4 KB per-warp staging area in shared memory:
//   in : item-major limbs of d = h - l (A rows, 32 B) and of l (the addend, as in fe_fold_fixed_f64_x2: column j starts at
//        2^52 + l_{j-1}, i.e. the bit pattern 0x43300000 : l_{j-1})      [2 folds x 32 items x (32 + 32) B]
//   out: the C fragments, read back item-major (16-byte chunks XOR-swizzled by the row: both directions conflict free)
// Fragments (PTX ISA, mma.m8n8k4.f64):  A (8x4) a0: row = lane >> 2 (item 8 mt + row), col = lane & 3 (half 4 ks + col)
//   B (4x8) b0: row = lane & 3 (half 4 ks + row), col = lane >> 2 (limb): four registers per lane hold the whole table
//   C (8x8) c0, c1: row = lane >> 2, cols 2 (lane & 3) and 2 (lane & 3) + 1
constexpr int kDmmaStageBytes = 4096;
struct DmmaTab {
    double b[4];  // this lane's B fragments, k-steps 0..3
};
__device__ __forceinline__ DmmaTab dmma_tab_load(const FixedMulF64& tab, int lane) {
    DmmaTab t;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
        t.b[ks] = tab.t[4 * ks + (lane & 3)][lane >> 2];
        asm volatile("" : "+d"(t.b[ks]));  // opaque: keep the fragment in registers instead of re-reading the constant bank per use
    }
    return t;
}
// Part 1: stage this lane's two pairs (l0, h0), (l1, h1); afterwards the four inputs are dead (the caller can issue its
// next loads) — only the top limbs of l0, l1 stay in registers.  ALL 32 lanes must call both parts (mma.sync, __syncwarp).
struct DmmaTop {
    uint32_t t0, t1;
};
template <class F>
__device__ __forceinline__ DmmaTop fe_fold_pair_dmma_stage(const Fe& l0, const Fe& l1, const Fe& h0, const Fe& h1, unsigned char* stage, int lane) {
    const Fe d0 = fe_sub<F>(h0, l0), d1 = fe_sub<F>(h1, l1);
    uint4* a0 = reinterpret_cast<uint4*>(stage + lane * 32);
    uint4* a1 = reinterpret_cast<uint4*>(stage + 1024 + lane * 32);
    uint4* b0 = reinterpret_cast<uint4*>(stage + 2048 + lane * 32);
    uint4* b1 = reinterpret_cast<uint4*>(stage + 3072 + lane * 32);
    a0[0] = make_uint4(d0.v[0], d0.v[1], d0.v[2], d0.v[3]);
    a0[1] = make_uint4(d0.v[4], d0.v[5], d0.v[6], d0.v[7]);
    a1[0] = make_uint4(d1.v[0], d1.v[1], d1.v[2], d1.v[3]);
    a1[1] = make_uint4(d1.v[4], d1.v[5], d1.v[6], d1.v[7]);
    b0[0] = make_uint4(l0.v[0], l0.v[1], l0.v[2], l0.v[3]);
    b0[1] = make_uint4(l0.v[4], l0.v[5], l0.v[6], l0.v[7]);
    b1[0] = make_uint4(l1.v[0], l1.v[1], l1.v[2], l1.v[3]);
    b1[1] = make_uint4(l1.v[4], l1.v[5], l1.v[6], l1.v[7]);
    __syncwarp();
    return DmmaTop{l0.v[7], l1.v[7]};
}
// Part 2: lo = l0 + r (h0 - l0), hi = l1 + r (h1 - l1) — the same field elements as fe_fold_fixed_f64_x2 / fe_fold_fixed.
template <class F>
__device__ __forceinline__ void fe_fold_pair_dmma_finish(Fe& lo, Fe& hi, const DmmaTop& top, const DmmaTab& tab, unsigned char* stage, int lane) {
    const int g = lane >> 2, c = lane & 3;
    double acc[2][4][2];
#pragma unroll
    for (int f = 0; f < 2; f++)
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            const uint32_t* lrow = reinterpret_cast<const uint32_t*>(stage + 2048 + f * 1024 + (8 * mt + g) * 32);
            const uint32_t la = c ? lrow[2 * c - 1] : 0u, lb = lrow[2 * c];
            acc[f][mt][0] = detail::f64_from_bits(0x43300000u, la);  // 2^52 + l_{j-1}: exact, the products leave 2^36 of headroom
            acc[f][mt][1] = detail::f64_from_bits(0x43300000u, lb);
        }
#pragma unroll
    for (int ks = 0; ks < 4; ks++)
#pragma unroll
        for (int f = 0; f < 2; f++)
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const unsigned short h = *reinterpret_cast<const unsigned short*>(stage + f * 1024 + (8 * mt + g) * 32 + (4 * ks + c) * 2);
                double a;
                asm("cvt.rn.f64.u16 %0, %1;" : "=d"(a) : "h"(h));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[f][mt][0]), "+d"(acc[f][mt][1])
                             : "d"(a), "d"(tab.b[ks]));
            }
    __syncwarp();  // every fragment load is done: the staging area is reused for the columns
    // C fragments -> item-major columns: fold f, item i at f * 2048 + i * 64, 16-byte chunk j stored at j ^ ((i >> 1) & 3)
#pragma unroll
    for (int f = 0; f < 2; f++)
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            const int row = 8 * mt + g;
            *reinterpret_cast<double2*>(stage + f * 2048 + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = make_double2(acc[f][mt][0], acc[f][mt][1]);
        }
    __syncwarp();
    double col[2][8];
#pragma unroll
    for (int f = 0; f < 2; f++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double2 v = *reinterpret_cast<const double2*>(stage + f * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
            col[f][2 * j] = v.x;
            col[f][2 * j + 1] = v.y;
        }
    __syncwarp();  // the next call's staging stores must not overtake these loads
    lo = f64_columns_reduce<F>(col[0], top.t0, true);
    hi = f64_columns_reduce<F>(col[1], top.t1, true);
}
template <class F>
__device__ __forceinline__ void fe_fold_pair_dmma(Fe& lo, Fe& hi, const Fe& l0, const Fe& l1, const Fe& h0, const Fe& h1,
                                                  const DmmaTab& tab, unsigned char* stage, int lane) {
    const DmmaTop top = fe_fold_pair_dmma_stage<F>(l0, l1, h0, h1, stage, lane);
    fe_fold_pair_dmma_finish<F>(lo, hi, top, tab, stage, lane);
}

}  // namespace zk
