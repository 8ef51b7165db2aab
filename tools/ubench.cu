// tools/ubench.cu — instruction-level throughput probes for the sm_100a integer pipe (stand-alone binary).
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench tools/ubench.cu
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>
#include "field29.cuh"
#include "../zk_b200/csrc/host_field.hpp"
using namespace zk;

constexpr int T = 256;
constexpr int ITERS = 2048;

// A: independent IMAD.WIDE.U32 (no carries)
__global__ void __launch_bounds__(T) kA(uint32_t seed, uint64_t* sink) {
    uint64_t a[8]; for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x + i;
    uint32_t x = seed * 2654435761u + threadIdx.x, y = x ^ 0x9e3779b9u;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(x), "r"(y));
    }
    uint64_t s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x1234567) sink[0] = s;
}
// B: cmad4 carry chains (4 x WIDE.X + addc), NCH independent accumulator sets
template <int NCH>
__global__ void __launch_bounds__(T) kB(uint32_t seed, uint32_t* sink) {
    uint32_t X[NCH][8], top[NCH];
    for (int c = 0; c < NCH; c++) { top[c] = 0; for (int i = 0; i < 8; i++) X[c][i] = seed + threadIdx.x + i + c; }
    uint32_t a0 = seed * 2654435761u + threadIdx.x, a2 = a0 ^ 0x9e3779b9u, a4 = a0 + 77, a6 = a2 + 99, b = a0 * 3 + 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) detail::cmad4(X[c], a0, a2, a4, a6, b, top[c]);
    }
    uint32_t s = 0; for (int c = 0; c < NCH; c++) { s ^= top[c]; for (int i = 0; i < 8; i++) s ^= X[c][i]; }
    if (s == 0x1234567) sink[0] = s;
}
// C: WIDE with carry-out only (pair mad.lo.cc + madc.hi.cc, then addc into a third register)
template <int NCH>
__global__ void __launch_bounds__(T) kC(uint32_t seed, uint32_t* sink) {
    uint32_t lo[NCH], hi[NCH], top[NCH];
    for (int c = 0; c < NCH; c++) { top[c] = 0; lo[c] = seed + c; hi[c] = seed * 3 + c; }
    uint32_t a0 = seed * 2654435761u + threadIdx.x, b = a0 * 3 + 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++)
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                             : "+r"(lo[c]), "+r"(hi[c]), "+r"(top[c]) : "r"(a0), "r"(b));
    }
    uint32_t s = 0; for (int c = 0; c < NCH; c++) s ^= top[c] ^ lo[c] ^ hi[c];
    if (s == 0x1234567) sink[0] = s;
}
// D: IADD3.X carry chains (8-limb add)
template <int NCH>
__global__ void __launch_bounds__(T) kD(uint32_t seed, uint32_t* sink) {
    uint32_t X[NCH][8];
    for (int c = 0; c < NCH; c++) for (int i = 0; i < 8; i++) X[c][i] = seed + threadIdx.x + i + c;
    uint32_t y = seed * 2654435761u + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++)
                asm volatile("add.cc.u32 %0,%0,%8;\n\taddc.cc.u32 %1,%1,%8;\n\taddc.cc.u32 %2,%2,%8;\n\taddc.cc.u32 %3,%3,%8;\n\t"
                             "addc.cc.u32 %4,%4,%8;\n\taddc.cc.u32 %5,%5,%8;\n\taddc.cc.u32 %6,%6,%8;\n\taddc.u32 %7,%7,%8;"
                             : "+r"(X[c][0]), "+r"(X[c][1]), "+r"(X[c][2]), "+r"(X[c][3]), "+r"(X[c][4]), "+r"(X[c][5]), "+r"(X[c][6]), "+r"(X[c][7]) : "r"(y));
    }
    uint32_t s = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 8; i++) s ^= X[c][i];
    if (s == 0x1234567) sink[0] = s;
}
// E: fe_mul chains with ILP
template <int ILP, bool LAZY>
__global__ void __launch_bounds__(T) kE(uint32_t seed, Fe* sink) {
    Fe x[ILP], y = fe_one<Fr381>(); y.v[0] ^= seed & 0xff;
    for (int k = 0; k < ILP; k++) { x[k] = fe_one<Fr381>(); x[k].v[0] ^= (seed + threadIdx.x + k) & 0xffff; }
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = LAZY ? fe_mul_lazy<Fr381>(y, x[k]) : fe_mul<Fr381>(x[k], y);
    }
    Fe s = x[0]; for (int k = 1; k < ILP; k++) for (int i = 0; i < 8; i++) s.v[i] ^= x[k].v[i];
    if (s.v[0] == 0x1234567 && s.v[7] == 0x7654321) sink[0] = s;
}


// F: independent DFMA chains
__global__ void __launch_bounds__(T) kF(uint32_t seed, double* sink) {
    double a[8]; for (int i = 0; i < 8; i++) a[i] = 1.0 + (seed + threadIdx.x + i) * 1e-9;
    double x = 1.0000001 + seed * 1e-12, y = 1e-13 * threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
    }
    double s = 0; for (int i = 0; i < 8; i++) s += a[i];
    if (s == 0.1234567) sink[0] = s;
}
// G: DFMA and IMAD.WIDE interleaved 1:1 (do the FP64 and FMA-int pipes overlap?)
__global__ void __launch_bounds__(T) kG(uint32_t seed, double* sink) {
    double a[4]; uint64_t c[4];
    for (int i = 0; i < 4; i++) { a[i] = 1.0 + (seed + threadIdx.x + i) * 1e-9; c[i] = seed + i; }
    double x = 1.0000001 + seed * 1e-12, y = 1e-13 * threadIdx.x;
    uint32_t p = seed * 2654435761u + threadIdx.x, q = p ^ 0x9e3779b9u;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i]) : "r"(p), "r"(q));
            }
    }
    double s = 0; for (int i = 0; i < 4; i++) s += a[i] + (double)c[i];
    if (s == 0.1234567) sink[0] = s;
}
// H: WIDE carry-in only (carry produced once by an add.cc outside, consumed by madc.lo; hi part without cc)
template <int NCH>
__global__ void __launch_bounds__(T) kH(uint32_t seed, uint32_t* sink) {
    uint32_t lo[NCH], hi[NCH], z[NCH];
    for (int c = 0; c < NCH; c++) { z[c] = seed + c + threadIdx.x; lo[c] = seed + c; hi[c] = seed * 3 + c; }
    uint32_t a0 = seed * 2654435761u + threadIdx.x, b = a0 * 3 + 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++)
                asm volatile("add.cc.u32 %2, %2, %3;\n\tmadc.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.u32 %1, %3, %4, %1;"
                             : "+r"(lo[c]), "+r"(hi[c]), "+r"(z[c]) : "r"(a0), "r"(b));
    }
    uint32_t s = 0; for (int c = 0; c < NCH; c++) s ^= z[c] ^ lo[c] ^ hi[c];
    if (s == 0x1234567) sink[0] = s;
}
// I: plain IADD3 three-input, independent
__global__ void __launch_bounds__(T) kI(uint32_t seed, uint32_t* sink) {
    uint32_t a[8]; for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x + i;
    uint32_t x = seed * 2654435761u + threadIdx.x, y = x ^ 0x12345;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(x), "r"(y));
    }
    uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x1234567) sink[0] = s;
}

// M: reduced-radix carry-free multiplier chains
template <int ILP>
__global__ void __launch_bounds__(T) kM(uint32_t seed, Fe* sink) {
    Fe29 x[ILP], y;
    for (int k = 0; k < 9; k++) y.l[k] = (seed * 77 + k * 1234567u) & kMask29;
    for (int c = 0; c < ILP; c++) for (int k = 0; k < 9; k++) x[c].l[k] = (seed + threadIdx.x * 9 + k + c * 31) & kMask29;
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int c = 0; c < ILP; c++) x[c] = mul29<Fr381>(x[c], y);
    }
    uint32_t s = 0; for (int c = 0; c < ILP; c++) for (int k = 0; k < 9; k++) s ^= x[c].l[k];
    if (s == 0x1234567) sink[0].v[0] = s;
}
// correctness: mul29(a, 32*b) must equal fe_mul(a, b) for Montgomery-form words
__device__ __forceinline__ uint64_t smix(uint64_t z) { z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31; return z; }
template <class F>
__global__ void kCheck(unsigned* bad, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fe a, b;
    for (int l = 0; l < 4; l++) {
        uint64_t u = smix(0x1234 + (uint64_t)i * 8 + l), v = smix(0x9876 + (uint64_t)i * 8 + l + 4);
        if (l == 3) { u &= 0x0FFFFFFFFFFFFFFFULL; v &= 0x0FFFFFFFFFFFFFFFULL; }
        a.v[2 * l] = (uint32_t)u; a.v[2 * l + 1] = (uint32_t)(u >> 32); b.v[2 * l] = (uint32_t)v; b.v[2 * l + 1] = (uint32_t)(v >> 32);
    }
    if (i == 0) { a = fe_zero<F>(); }
    if (i == 1) { for (int k = 0; k < 8; k++) { a.v[k] = F::p(k); b.v[k] = F::p(k); } a.v[0] -= 1; b.v[0] -= 1; }  // p-1
    Fe expect = fe_mul<F>(a, b);
    Fe b32 = b;
    for (int d = 0; d < 5; d++) b32 = fe_add<F>(b32, b32);
    Fe29 y = mul29<F>(unpack29(a), unpack29(b32));
    Fe got = fe_reduce_once<F>(pack29(y));
    bool same = true;
    for (int k = 0; k < 8; k++) same &= (got.v[k] == expect.v[k]);
    // round trip of the regrouping
    Fe rt = pack29(unpack29(a));
    for (int k = 0; k < 8; k++) same &= (rt.v[k] == a.v[k]);
    if (!same) atomicAdd(bad, 1u);
}

// W1: pure 32x32->64 products, operands data dependent (cannot be hoisted): IMAD.WIDE.U32 R, Rlo, Rhi, RZ
__global__ void __launch_bounds__(T) kW1(uint32_t seed, uint64_t* sink) {
    uint64_t a[8]; for (int i = 0; i < 8; i++) a[i] = ((uint64_t)(seed * 77 + i) << 32) | (seed + threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, hi;}" : "+l"(a[i]));
    }
    uint64_t s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x1234567) sink[0] = s;
}
// W2: accumulate form with a data-dependent multiplicand: mad.wide.u32 acc, acc.lo, y, acc
__global__ void __launch_bounds__(T) kW2(uint32_t seed, uint64_t* sink) {
    uint64_t a[8]; for (int i = 0; i < 8; i++) a[i] = ((uint64_t)(seed * 77 + i) << 32) | (seed + threadIdx.x + i);
    uint32_t y = seed * 2654435761u + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(a[i]) : "r"(y));
    }
    uint64_t s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x1234567) sink[0] = s;
}
// W3: 32-bit IMAD (lo) with data-dependent multiplicand
__global__ void __launch_bounds__(T) kW3(uint32_t seed, uint32_t* sink) {
    uint32_t a[8]; for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x + i;
    uint32_t y = seed * 2654435761u + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(y));
    }
    uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x1234567) sink[0] = s;
}
// W4: IMAD.HI with data-dependent multiplicand
__global__ void __launch_bounds__(T) kW4(uint32_t seed, uint32_t* sink) {
    uint32_t a[8]; for (int i = 0; i < 8; i++) a[i] = seed * 0x9e3779b9u + threadIdx.x + i;
    uint32_t y = seed * 2654435761u + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(y));
    }
    uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x1234567) sink[0] = s;
}
// W5: FFMA for reference (fp32 pipe = fmaheavy + fmalite)
__global__ void __launch_bounds__(T) kW5(uint32_t seed, float* sink) {
    float a[8]; for (int i = 0; i < 8; i++) a[i] = 1.0f + (seed + threadIdx.x + i) * 1e-6f;
    float y = 1.0f + seed * 1e-7f;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(y));
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    if (s == 0.1234567f) sink[0] = s;
}

// X: fixed-multiplier product vs the general multiplier
template <class F>
__global__ void kCheckFixed(unsigned* bad, int n, FixedMul tab, Fe r_mont) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fe a;
    for (int l = 0; l < 4; l++) {
        uint64_t u = smix(0x4321 + (uint64_t)i * 4 + l);
        if (l == 3) u &= 0x0FFFFFFFFFFFFFFFULL;
        a.v[2 * l] = (uint32_t)u; a.v[2 * l + 1] = (uint32_t)(u >> 32);
    }
    if (i == 0) a = fe_zero<F>();
    if (i == 1) { for (int k = 0; k < 8; k++) a.v[k] = F::p(k); a.v[0] -= 1; }
    if (i == 2) { a = fe_zero<F>(); a.v[0] = 0xffffffffu; }
    if (i == 3) { for (int k = 0; k < 8; k++) a.v[k] = 0xffffffffu; a.v[7] = F::p(7) - 1; }
    Fe expect = fe_mul<F>(a, r_mont), got = fe_mul_fixed<F>(a, tab);
    bool same = true;
    for (int k = 0; k < 8; k++) same &= (got.v[k] == expect.v[k]);
    if (!same) atomicAdd(bad, 1u);
}
template <int ILP>
__global__ void __launch_bounds__(T) kFx(uint32_t seed, Fe* sink, FixedMul tab) {
    Fe x[ILP];
    for (int k = 0; k < ILP; k++) { x[k] = fe_one<Fr381>(); x[k].v[0] ^= (seed + threadIdx.x + k) & 0xffff; }
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = fe_mul_fixed<Fr381>(x[k], tab);
    }
    Fe s = x[0]; for (int k = 1; k < ILP; k++) for (int i = 0; i < 8; i++) s.v[i] ^= x[k].v[i];
    if (s.v[0] == 0x1234567 && s.v[7] == 0x7654321) sink[0] = s;
}

template <class L> float run(L&& launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 1 && ms < best) best = ms; }
    return best;
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    void* sink; cudaMalloc(&sink, 256);
    const double sm_clk = 1.965e9 * sms;  // report per SM-clock at the max clock (nvidia-smi shows 1965 MHz under this load)
    {
        unsigned* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
        kCheck<Fr381><<<4096, 256>>>(bad, 1 << 20);
        unsigned h381 = 0; cudaMemcpy(&h381, bad, 4, cudaMemcpyDeviceToHost); cudaMemset(bad, 0, 4);
        kCheck<Fr377><<<4096, 256>>>(bad, 1 << 20);
        unsigned h377 = 0; cudaMemcpy(&h377, bad, 4, cudaMemcpyDeviceToHost);
        printf("mul29 vs fe_mul mismatches over 2^20 pairs: Fr381 %u, Fr377 %u  (%s)\n", h381, h377, cudaGetErrorString(cudaGetLastError()));
    }
    FixedMul tab381;
    for (int fid = 0; fid < 2; fid++) {
        zk::host::Field HF(fid);
        zk::host::El r = HF.from_u64(0x123456789abcdefULL);
        for (int k = 0; k < 5; k++) r = HF.mul(r, HF.add(r, HF.from_u64(77 + k)));  // some dense element
        FixedMul tab; zk::host::fixed_mul_table(HF, r, tab.v);
        Fe rm; memcpy(rm.v, r.v, 32);
        unsigned* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
        if (fid == 0) { kCheckFixed<Fr381><<<4096, 256>>>(bad, 1 << 20, tab, rm); tab381 = tab; }
        else kCheckFixed<Fr377><<<4096, 256>>>(bad, 1 << 20, tab, rm);
        unsigned h = 0; cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
        printf("fe_mul_fixed vs fe_mul mismatches over 2^20 values, field %d: %u (%s)\n", fid, h, cudaGetErrorString(cudaGetLastError()));
    }
    for (int bps : {4, 8}) {
        const int blocks = sms * bps; const double thr = (double)blocks * T;
        printf("--- %d blocks/SM x %d threads (%d warps/SMSP)\n", bps, T, bps * T / 32 / 4);
        float ms = run([&] { kA<<<blocks, T>>>(7, (uint64_t*)sink); });
        printf("A  IMAD.WIDE independent      : %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kW1<<<blocks, T>>>(7, (uint64_t*)sink); });
        printf("W1 IMAD.WIDE product (dep ops): %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kW2<<<blocks, T>>>(7, (uint64_t*)sink); });
        printf("W2 mad.wide accumulate (dep)  : %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kW3<<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("W3 IMAD lo (dep)              : %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kW4<<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("W4 IMAD.HI (dep)              : %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kW5<<<blocks, T>>>(7, (float*)sink); });
        printf("W5 FFMA (dep)                 : %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kB<1><<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("B1 cmad4 chain x1 (4W.X+addc) : %7.2f WIDE/clk/SM\n", thr * ITERS * 2 * 1 * 4 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kB<2><<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("B2 cmad4 chain x2             : %7.2f WIDE/clk/SM\n", thr * ITERS * 2 * 2 * 4 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kB<4><<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("B4 cmad4 chain x4             : %7.2f WIDE/clk/SM\n", thr * ITERS * 2 * 4 * 4 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kC<8><<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("C8 WIDE carry-out + addc      : %7.2f WIDE/clk/SM\n", thr * ITERS * 4 * 8 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kD<4><<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("D4 IADD3.X 8-limb chains      : %7.2f IADD/clk/SM\n", thr * ITERS * 2 * 4 * 8 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kF<<<blocks, T>>>(7, (double*)sink); });
        printf("F  DFMA independent           : %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kG<<<blocks, T>>>(7, (double*)sink); });
        printf("G  DFMA+IMAD.WIDE 1:1         : %7.2f instr/clk/SM (both kinds counted)\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kH<8><<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("H8 add.cc + WIDE carry-in     : %7.2f WIDE/clk/SM\n", thr * ITERS * 4 * 8 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kI<<<blocks, T>>>(7, (uint32_t*)sink); });
        printf("I  IADD3 (3-input) independent: %7.2f instr/clk/SM\n", thr * ITERS * 32 / (ms * 1e-3) / sm_clk);
        ms = run([&] { kM<1><<<blocks, T>>>(7, (Fe*)sink); });
        printf("M1 mul29 ILP1                 : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 1 / (ms * 1e-3) / sm_clk, thr * 256 * 1 / (ms * 1e-3));
        ms = run([&] { kM<2><<<blocks, T>>>(7, (Fe*)sink); });
        printf("M2 mul29 ILP2                 : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 2 / (ms * 1e-3) / sm_clk, thr * 256 * 2 / (ms * 1e-3));
        ms = run([&] { kFx<1><<<blocks, T>>>(7, (Fe*)sink, tab381); });
        printf("X1 fe_mul_fixed ILP1          : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 1 / (ms * 1e-3) / sm_clk, thr * 256 * 1 / (ms * 1e-3));
        ms = run([&] { kFx<2><<<blocks, T>>>(7, (Fe*)sink, tab381); });
        printf("X2 fe_mul_fixed ILP2          : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 2 / (ms * 1e-3) / sm_clk, thr * 256 * 2 / (ms * 1e-3));
        ms = run([&] { kE<1, false><<<blocks, T>>>(7, (Fe*)sink); });
        printf("E1 fe_mul ILP1                : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 1 / (ms * 1e-3) / sm_clk, thr * 256 * 1 / (ms * 1e-3));
        ms = run([&] { kE<2, false><<<blocks, T>>>(7, (Fe*)sink); });
        printf("E2 fe_mul ILP2                : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 2 / (ms * 1e-3) / sm_clk, thr * 256 * 2 / (ms * 1e-3));
        ms = run([&] { kE<4, false><<<blocks, T>>>(7, (Fe*)sink); });
        printf("E4 fe_mul ILP4                : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 4 / (ms * 1e-3) / sm_clk, thr * 256 * 4 / (ms * 1e-3));
        ms = run([&] { kE<4, true><<<blocks, T>>>(7, (Fe*)sink); });
        printf("E4L fe_mul_lazy ILP4          : %7.3f mul/clk/SM  (%.3e mul/s)\n", thr * 256 * 4 / (ms * 1e-3) / sm_clk, thr * 256 * 4 / (ms * 1e-3));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
