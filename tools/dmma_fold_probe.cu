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

#include "fold_dmma.cuh"

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
