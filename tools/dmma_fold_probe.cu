// dmma_fold_probe.cu — stand-alone throughput of the two folds of an item by a launch-wide challenge, at the round kernel's
// occupancy (128 threads, 4 blocks per SM), register-resident operands:
//   int   : fe_fold_fixed x2 (76 IMAD.WIDE each)         dfma : fe_fold_fixed_f64_x2 (256 DFMA + 128 uniform loads)
//   dmma  : fe_fold_pair_dmma (32 DMMA.8x8x4 per warp through a 4 KB shared-memory staging area)
// and a bit-exact cross-check of the three.  The DMMA fold lives HERE, not in the product: it is the fastest of the three
// stand-alone (B200: 9.5e10 folds/s without the addend trick, 7.8e10 with it, against 6.2e10 DFMA and 8.6e10 IMAD.WIDE)
// and leaves the integer pipe free, but both ways of putting it into the fused round kernel measured SLOWER than the DFMA
// folds (2.92 and 3.19 ms against 2.44 ms for the first fused step of the 2^26 degree-3 proof): next to the
// running products, the prefetched quadruple and the 34 KB of wide accumulators there are neither registers nor shared
// memory left for its fragments and staging area — ptxas serialises the DMMA chains and spills, or the products' operands
// have to be re-read from L2.  Kept as a measured data point and a starting point for a kernel built around it.  build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I zk_b200/csrc
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "field_f64.cuh"
#include "host_field.hpp"

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

using namespace zk;
constexpr int kIters = 256;
constexpr int kThreadsP = 128;

template <class F, int MODE>
__global__ void __launch_bounds__(kThreadsP, 4) fold_probe(Fe* out, uint32_t seed, const __grid_constant__ FixedMul tab,
                                                          const __grid_constant__ FixedMulF64Sel tab64, int check) {
    extern __shared__ unsigned char stage_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = stage_all + warp * kDmmaStageBytes;
    Fe x0 = fe_one<F>(), x1 = fe_one<F>(), x2 = fe_one<F>(), x3 = fe_one<F>();
    x0.v[0] ^= (seed + threadIdx.x * 7 + blockIdx.x) & 0xffffff; x1.v[1] ^= seed * 3 + threadIdx.x; x2.v[2] ^= seed + 5 * threadIdx.x; x3.v[0] ^= 77 + threadIdx.x;
    x0 = fe_reduce_once<F>(x0); x1 = fe_reduce_once<F>(x1); x2 = fe_reduce_once<F>(x2); x3 = fe_reduce_once<F>(x3);
    DmmaTab dt{};
    if (MODE == 2) dt = dmma_tab_load(tab64.t[0], lane);
#pragma unroll 1
    for (int i = 0; i < kIters; i++) {
        Fe lo, hi;
        if (MODE == 0) { lo = fe_fold_fixed<F>(x0, x2, tab); hi = fe_fold_fixed<F>(x1, x3, tab); }
        else if (MODE == 1) fe_fold_fixed_f64_x2<F>(lo, hi, x0, x1, x2, x3, tab64.t[(i * seed) >> 30]);
        else fe_fold_pair_dmma<F>(lo, hi, x0, x1, x2, x3, dt, stage, lane);
        x2 = x0; x3 = x1; x0 = lo; x1 = hi;
    }
    if (check || (x0.v[0] == 0x1234567 && x1.v[7] == 0x7654321)) { out[2 * (blockIdx.x * kThreadsP + threadIdx.x)] = x0; out[2 * (blockIdx.x * kThreadsP + threadIdx.x) + 1] = x1; }
}

template <class F>
int run(int field) {
    host::Field HF(field);
    host::El r = HF.from_u64(0x123456789abcdefull);
    r = HF.mul(r, HF.mul(r, r));
    FixedMul tab;
    FixedMulF64Sel tab64;
    host::fixed_mul_table(HF, r, tab.v);
    host::fixed_mul_table_f64(HF, r, tab64.t[0].t);
    tab64.t[1] = tab64.t[0];
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4;
    Fe* out[3];
    for (int m = 0; m < 3; m++) cudaMalloc(&out[m], (size_t)grid * kThreadsP * 2 * sizeof(Fe));
    cudaFuncSetAttribute(fold_probe<F, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kDmmaStageBytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[3] = {"int  (fe_fold_fixed x2)", "dfma (fe_fold_fixed_f64_x2)", "dmma (fe_fold_pair_dmma)"};
    for (int m = 0; m < 3; m++) {
        float ms = 0;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (m == 0) fold_probe<F, 0><<<grid, kThreadsP>>>(out[0], 12345, tab, tab64, 1);
            if (m == 1) fold_probe<F, 1><<<grid, kThreadsP>>>(out[1], 12345, tab, tab64, 1);
            if (m == 2) fold_probe<F, 2><<<grid, kThreadsP, 4 * kDmmaStageBytes>>>(out[2], 12345, tab, tab64, 1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double folds = (double)grid * kThreadsP * kIters * 2;
        printf("field %d %-30s %8.3f ms  %.3e folds/s  (%s)\n", field, names[m], ms, folds / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    }
    const size_t n = (size_t)grid * kThreadsP * 2;
    Fe* h[3];
    for (int m = 0; m < 3; m++) { h[m] = new Fe[n]; cudaMemcpy(h[m], out[m], n * sizeof(Fe), cudaMemcpyDeviceToHost); }
    long bad1 = 0, bad2 = 0;
    for (size_t i = 0; i < n; i++) { bad1 += std::memcmp(&h[0][i], &h[1][i], 32) != 0; bad2 += std::memcmp(&h[0][i], &h[2][i], 32) != 0; }
    printf("field %d cross-check over %zu results: dfma vs int %ld mismatches, dmma vs int %ld mismatches\n", field, n, bad1, bad2);
    return (bad1 || bad2) ? 1 : 0;
}

int main() { return run<Fr381>(0) | run<Fr377>(1); }
