// imma_fold_probe.cu — stand-alone throughput of the fold of an item pair by a launch-wide challenge, at the round kernel's
// occupancy (128 threads, 4 blocks per SM), register-resident operands:
//   int  : fe_fold_fixed x2 (76 IMAD.WIDE each)         dfma : fe_fold_fixed_f64_x2 (256 DFMA + 128 uniform loads)
//   imma : fe_fold_imma x2 (8 IMMA.16832.U8.U8 per warp and fold through a 2.5 KB shared-memory staging area)
// a bit-exact cross-check of the three, and the raw issue rate of the IMMA instruction.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I zk_b200/csrc -I include tools/imma_fold_probe.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "field_f64.cuh"
#include "host_field.hpp"

#include "fold_imma.cuh"

using namespace zk;
constexpr int kIters = 256;
constexpr int kThreadsP = 128;

template <class F, int MODE>
__global__ void __launch_bounds__(kThreadsP, 4) fold_probe(Fe* out, uint32_t seed, const __grid_constant__ FixedMul tab,
                                                          const __grid_constant__ FixedMulF64Sel tab64, const FixedMulI8* tab8, int check) {
    extern __shared__ __align__(16) unsigned char stage_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = stage_all + warp * 2 * kImmaStageBytes;
    Fe x0 = fe_one<F>(), x1 = fe_one<F>(), x2 = fe_one<F>(), x3 = fe_one<F>();
    x0.v[0] ^= (seed + threadIdx.x * 7 + blockIdx.x) & 0xffffff; x1.v[1] ^= seed * 3 + threadIdx.x; x2.v[2] ^= seed + 5 * threadIdx.x; x3.v[0] ^= 77 + threadIdx.x;
    x0 = fe_reduce_once<F>(x0); x1 = fe_reduce_once<F>(x1); x2 = fe_reduce_once<F>(x2); x3 = fe_reduce_once<F>(x3);
    ImmaTab it{};
    if (MODE >= 2) it = imma_tab_load(*tab8, lane);
#pragma unroll 1
    for (int i = 0; i < kIters; i++) {
        Fe lo, hi;
        if (MODE == 0) { lo = fe_fold_fixed<F>(x0, x2, tab); hi = fe_fold_fixed<F>(x1, x3, tab); }
        else if (MODE == 1) fe_fold_fixed_f64_x2<F>(lo, hi, x0, x1, x2, x3, tab64.t[(i * seed) >> 30]);
        else if (MODE == 2) { lo = fe_fold_imma<F>(x0, x2, it, stage, lane); hi = fe_fold_imma<F>(x1, x3, it, stage, lane); }
        else { Fe o[2]; const Fe l2[2] = {x0, x1}, h2[2] = {x2, x3}; fe_fold_imma_n<F, 2>(o, l2, h2, it, stage, lane); lo = o[0]; hi = o[1]; }
        x2 = x0; x3 = x1; x0 = lo; x1 = hi;
    }
    if (check || (x0.v[0] == 0x1234567 && x1.v[7] == 0x7654321)) { out[2 * (blockIdx.x * kThreadsP + threadIdx.x)] = x0; out[2 * (blockIdx.x * kThreadsP + threadIdx.x) + 1] = x1; }
}

// raw instruction rate: four independent accumulator sets per warp, operands loop-carried
__global__ void __launch_bounds__(256) imma_rate(int* out, uint32_t seed, int iters) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 ^ 0x55aa, b1 = a0 ^ 0xaa55;
    int c[4][4] = {};
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) s += c[u][0] ^ c[u][1] ^ c[u][2] ^ c[u][3];
    if (s == 0x1234567) out[0] = s;
}

template <class F>
int run(int field) {
    host::Field HF(field);
    host::El r = HF.from_u64(0x123456789abcdefull);
    r = HF.mul(r, HF.mul(r, r));
    FixedMul tab;
    FixedMulF64Sel tab64;
    FixedMulI8 tab8, *d_tab8 = nullptr;
    host::fixed_mul_table(HF, r, tab.v);
    host::fixed_mul_table_f64(HF, r, tab64.t[0].t);
    tab64.t[1] = tab64.t[0];
    fixed_mul_table_i8(HF, r, &tab8);
    cudaMalloc(&d_tab8, sizeof(tab8));
    cudaMemcpy(d_tab8, &tab8, sizeof(tab8), cudaMemcpyHostToDevice);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4;
    Fe* out[4];
    for (int m = 0; m < 4; m++) cudaMalloc(&out[m], (size_t)grid * kThreadsP * 2 * sizeof(Fe));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"int  (fe_fold_fixed x2)", "dfma (fe_fold_fixed_f64_x2)", "imma (fe_fold_imma x2)", "imma (fe_fold_imma_n<2>)"};
    for (int m = 0; m < 4; m++) {
        float ms = 0;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (m == 0) fold_probe<F, 0><<<grid, kThreadsP>>>(out[0], 12345, tab, tab64, d_tab8, 1);
            if (m == 1) fold_probe<F, 1><<<grid, kThreadsP>>>(out[1], 12345, tab, tab64, d_tab8, 1);
            if (m == 2) fold_probe<F, 2><<<grid, kThreadsP, 8 * kImmaStageBytes>>>(out[2], 12345, tab, tab64, d_tab8, 1);
            if (m == 3) fold_probe<F, 3><<<grid, kThreadsP, 8 * kImmaStageBytes>>>(out[3], 12345, tab, tab64, d_tab8, 1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double folds = (double)grid * kThreadsP * kIters * 2;
        printf("field %d %-30s %8.3f ms  %.3e folds/s  (%s)\n", field, names[m], ms, folds / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    }
    const size_t n = (size_t)grid * kThreadsP * 2;
    Fe* h[4];
    for (int m = 0; m < 4; m++) { h[m] = new Fe[n]; cudaMemcpy(h[m], out[m], n * sizeof(Fe), cudaMemcpyDeviceToHost); }
    long bad1 = 0, bad2 = 0;
    for (size_t i = 0; i < n; i++) { bad1 += std::memcmp(&h[0][i], &h[1][i], 32) != 0; bad2 += std::memcmp(&h[0][i], &h[2][i], 32) != 0; bad2 += std::memcmp(&h[0][i], &h[3][i], 32) != 0; }
    printf("field %d cross-check over %zu results: dfma vs int %ld mismatches, imma (both forms) vs int %ld mismatches\n", field, n, bad1, bad2);
    if (bad2) {
        for (size_t i = 0, shown = 0; i < n && shown < 4; i++)
            if (std::memcmp(&h[0][i], &h[2][i], 32)) {
                printf("  [%zu] int ", i); for (int k = 7; k >= 0; k--) printf("%08x", h[0][i].v[k]);
                printf("\n       imma "); for (int k = 7; k >= 0; k--) printf("%08x", h[2][i].v[k]);
                printf("\n"); shown++;
            }
    }
    return (bad1 || bad2) ? 1 : 0;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int* d = nullptr;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bps = 1; bps <= 2; bps++) {
        float ms = 0;
        const int iters = 1 << 16;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            imma_rate<<<sms * bps, 256>>>(d, 99, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double n = (double)sms * bps * 8 * iters * 4;
        printf("IMMA.16832.U8.U8: %d blocks of 8 warps per SM: %.3e instr/s = %.1f per SM per 1000 clocks at 1965 MHz, %.3e MAC/s\n", bps, n / (ms * 1e-3), n / (ms * 1e-3) / sms / 1.965e9 * 1000, n / (ms * 1e-3) * 4096);
    }
    return run<Fr381>(0) | run<Fr377>(1);
}
