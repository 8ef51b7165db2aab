// microbench.cu — measures the roofline denominators the path is judged against, on the box itself:
// IMAD.WIDE.U32 / IMAD / IADD3 issue rates, stand-alone Montgomery-multiplication throughput (the
// field-mul ceiling of this multiplier) and 256-bit streaming bandwidth.  SURVEY.md 8d asks for the
// integer peak to be measured rather than assumed; MEASURED_PEAKS.json only carries HBM and bf16.
#include "kernels.h"
#include "field_f64.cuh"
#include "host_field.hpp"

namespace zk {
namespace {

constexpr int kThreads = 256;
constexpr int kIters = 4096;

__global__ void __launch_bounds__(kThreads) imad_wide_kernel(uint32_t seed, uint64_t* sink, long long* clocks) {
    uint64_t a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint32_t x = seed * 2654435761u + threadIdx.x, y = x ^ 0x9e3779b9u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile(
                // operands are the halves of the running value, so nothing is loop invariant (a loop-invariant
                // product is hoisted by ptxas and the probe then measures 64-bit adds, not multiplies)
                "{.reg .u32 l, h;\n\t"
                "mov.b64 {l,h}, %0; mul.wide.u32 %0, l, h;\n\tmov.b64 {l,h}, %1; mul.wide.u32 %1, l, h;\n\t"
                "mov.b64 {l,h}, %2; mul.wide.u32 %2, l, h;\n\tmov.b64 {l,h}, %3; mul.wide.u32 %3, l, h;\n\t"
                "mov.b64 {l,h}, %4; mul.wide.u32 %4, l, h;\n\tmov.b64 {l,h}, %5; mul.wide.u32 %5, l, h;\n\t"
                "mov.b64 {l,h}, %6; mul.wide.u32 %6, l, h;\n\tmov.b64 {l,h}, %7; mul.wide.u32 %7, l, h;}"
                : "+l"(a0), "+l"(a1), "+l"(a2), "+l"(a3), "+l"(a4), "+l"(a5), "+l"(a6), "+l"(a7)
                : "r"(x), "r"(y));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *clocks = t1 - t0;
    uint64_t s = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (s == 0x1234567) sink[0] = s;
}
__global__ void __launch_bounds__(kThreads) imad_lo_kernel(uint32_t seed, uint32_t* sink) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint32_t x = seed * 2654435761u + threadIdx.x, y = x ^ 0x9e3779b9u;
#pragma unroll 1
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile(
                "mad.lo.u32 %0, %0, %9, %0;\n\tmad.lo.u32 %1, %1, %9, %1;\n\tmad.lo.u32 %2, %2, %9, %2;\n\t"
                "mad.lo.u32 %3, %3, %9, %3;\n\tmad.lo.u32 %4, %4, %9, %4;\n\tmad.lo.u32 %5, %5, %9, %5;\n\t"
                "mad.lo.u32 %6, %6, %9, %6;\n\tmad.lo.u32 %7, %7, %9, %7;"
                : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7)
                : "r"(x), "r"(y));
        }
    }
    uint32_t s = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (s == 0x1234567) sink[0] = s;
}
__global__ void __launch_bounds__(kThreads) iadd3_kernel(uint32_t seed, uint32_t* sink) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint32_t x = seed * 2654435761u + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile(
                "add.u32 %0, %0, %8;\n\tadd.u32 %1, %1, %8;\n\tadd.u32 %2, %2, %8;\n\tadd.u32 %3, %3, %8;\n\t"
                "add.u32 %4, %4, %8;\n\tadd.u32 %5, %5, %8;\n\tadd.u32 %6, %6, %8;\n\tadd.u32 %7, %7, %8;"
                : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7)
                : "r"(x));
        }
    }
    uint32_t s = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (s == 0x1234567) sink[0] = s;
}
__global__ void __launch_bounds__(kThreads) mixed_kernel(uint32_t seed, uint64_t* sink) {
    uint64_t a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    uint32_t b0 = seed, b1 = seed + 1, b2 = seed + 2, b3 = seed + 3;
    uint32_t x = seed * 2654435761u + threadIdx.x, y = x ^ 0x9e3779b9u;
#pragma unroll 1
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile(
                "{.reg .u32 l, h;\n\t"
                "mov.b64 {l,h}, %0; mul.wide.u32 %0, l, h;\n\tadd.u32 %4, %4, %8;\n\tmov.b64 {l,h}, %1; mul.wide.u32 %1, l, h;\n\tadd.u32 %5, %5, %8;\n\t"
                "mov.b64 {l,h}, %2; mul.wide.u32 %2, l, h;\n\tadd.u32 %6, %6, %8;\n\tmov.b64 {l,h}, %3; mul.wide.u32 %3, l, h;\n\tadd.u32 %7, %7, %8;}"
                : "+l"(a0), "+l"(a1), "+l"(a2), "+l"(a3), "+r"(b0), "+r"(b1), "+r"(b2), "+r"(b3)
                : "r"(x), "r"(y));
        }
    }
    uint64_t s = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
    if (s == 0x1234567) sink[0] = s;
}
__global__ void __launch_bounds__(kThreads) dfma_kernel(uint32_t seed, double* sink) {
    double a0 = 1.0 + seed * 1e-9 + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double x = 1.0000001 + seed * 1e-12, y = 1e-13 * threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile(
                "fma.rn.f64 %0, %0, %8, %9;\n\tfma.rn.f64 %1, %1, %8, %9;\n\tfma.rn.f64 %2, %2, %8, %9;\n\t"
                "fma.rn.f64 %3, %3, %8, %9;\n\tfma.rn.f64 %4, %4, %8, %9;\n\tfma.rn.f64 %5, %5, %8, %9;\n\t"
                "fma.rn.f64 %6, %6, %8, %9;\n\tfma.rn.f64 %7, %7, %8, %9;"
                : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7)
                : "d"(x), "d"(y));
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 0.1234567) sink[0] = s;
}
constexpr int kMulIters = 256;
template <class F>
__global__ void __launch_bounds__(kThreads) fe_mul_fixed_f64_kernel(uint32_t seed, Fe* sink, const __grid_constant__ FixedMulF64Sel tab) {
    Fe x[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        x[k] = fe_one<F>();
        x[k].v[0] ^= (seed + threadIdx.x + k) & 0xffff;
    }
#pragma unroll 1
    for (int i = 0; i < kMulIters; i++) fe_mul_fixed_f64_x2<F>(x[0], x[1], tab.t[(i * seed) >> 30]);  // index: always 0, loop-variant for ptxas
    Fe s = fe_add<F>(x[0], x[1]);
    if (s.v[0] == 0x1234567 && s.v[7] == 0x7654321) sink[0] = s;
}
template <class F>
__global__ void __launch_bounds__(kThreads) fe_mul_kernel(uint32_t seed, Fe* sink) {
    Fe x[4], y;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = fe_one<F>();
        x[k].v[0] ^= (seed + threadIdx.x + k) & 0xffff;
    }
    y = fe_one<F>();
    y.v[0] ^= seed & 0xff;
#pragma unroll 1
    for (int i = 0; i < kMulIters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] = fe_mul<F>(x[k], y);
    }
    Fe s = fe_add<F>(fe_add<F>(x[0], x[1]), fe_add<F>(x[2], x[3]));
    if (s.v[0] == 0x1234567 && s.v[7] == 0x7654321) sink[0] = s;
}
__global__ void __launch_bounds__(kThreads) copy_kernel(const Fe* in, Fe* out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) st_fe(out + j, ld_fe_stream(in + j));
}
__global__ void __launch_bounds__(kThreads) read_kernel(const Fe* in, uint64_t n, Fe* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    uint32_t acc = 0;
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
        Fe x = ld_fe_stream(in + j);
        acc ^= x.v[0] ^ x.v[3] ^ x.v[4] ^ x.v[7] ^ x.v[1] ^ x.v[2] ^ x.v[5] ^ x.v[6];
    }
    if (acc == 0x1234567) sink[0].v[0] = acc;
}

template <class L>
cudaError_t time_ms(L&& launch, cudaStream_t st, int reps, float* best) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    *best = 1e30f;
    cudaError_t err = cudaSuccess;
    for (int r = 0; r < reps + 2 && err == cudaSuccess; r++) {
        cudaEventRecord(e0, st);
        launch();
        cudaEventRecord(e1, st);
        err = cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < *best) *best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (err == cudaSuccess) err = cudaGetLastError();
    return err;
}
}  // namespace

cudaError_t run_microbench(int field, MicrobenchResult* out, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8;
    void* sink = nullptr;
    long long* clk = nullptr;
    cudaError_t e = cudaMalloc(&sink, 256);
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void**)&clk, 8);
    if (e != cudaSuccess) return e;
    float ms;
    const double threads = (double)blocks * kThreads;
    e = time_ms([&] { imad_wide_kernel<<<blocks, kThreads, 0, st>>>(7u, (uint64_t*)sink, clk); }, st, 5, &ms);
    if (e != cudaSuccess) return e;
    out->imad_wide_per_s = threads * kIters * 32.0 / (ms * 1e-3);
    long long cyc = 0;
    cudaMemcpy(&cyc, clk, 8, cudaMemcpyDeviceToHost);
    out->sm_clock_mhz = (double)cyc / (ms * 1e-3) / 1e6;  // lower bound: block 0's active window over the whole launch
    e = time_ms([&] { imad_lo_kernel<<<blocks, kThreads, 0, st>>>(7u, (uint32_t*)sink); }, st, 5, &ms);
    if (e != cudaSuccess) return e;
    out->imad_lo_per_s = threads * kIters * 32.0 / (ms * 1e-3);
    e = time_ms([&] { iadd3_kernel<<<blocks, kThreads, 0, st>>>(7u, (uint32_t*)sink); }, st, 5, &ms);
    if (e != cudaSuccess) return e;
    out->iadd3_per_s = threads * kIters * 32.0 / (ms * 1e-3);
    e = time_ms([&] { mixed_kernel<<<blocks, kThreads, 0, st>>>(7u, (uint64_t*)sink); }, st, 5, &ms);
    if (e != cudaSuccess) return e;
    out->mixed_per_s = threads * kIters * 32.0 / (ms * 1e-3);
    if (field == Fr381::ID)
        e = time_ms([&] { fe_mul_kernel<Fr381><<<blocks, kThreads, 0, st>>>(7u, (Fe*)sink); }, st, 5, &ms);
    else
        e = time_ms([&] { fe_mul_kernel<Fr377><<<blocks, kThreads, 0, st>>>(7u, (Fe*)sink); }, st, 5, &ms);
    if (e != cudaSuccess) return e;
    out->fe_mul_per_s = threads * kMulIters * 4.0 / (ms * 1e-3);
    e = time_ms([&] { dfma_kernel<<<blocks, kThreads, 0, st>>>(7u, (double*)sink); }, st, 5, &ms);
    if (e != cudaSuccess) return e;
    out->dfma_per_s = threads * kIters * 32.0 / (ms * 1e-3);
    {
        host::Field HF(field);
        host::El r = HF.from_u64(0x123456789abcdefULL);
        for (int k = 0; k < 5; k++) r = HF.mul(r, HF.add(r, HF.from_u64(77 + k)));  // some dense element
        FixedMulF64Sel tab;
        host::fixed_mul_table_f64(HF, r, tab.t[0].t);
        tab.t[1] = tab.t[0];
        if (field == Fr381::ID)
            e = time_ms([&] { fe_mul_fixed_f64_kernel<Fr381><<<blocks, kThreads, 0, st>>>(7u, (Fe*)sink, tab); }, st, 5, &ms);
        else
            e = time_ms([&] { fe_mul_fixed_f64_kernel<Fr377><<<blocks, kThreads, 0, st>>>(7u, (Fe*)sink, tab); }, st, 5, &ms);
        if (e != cudaSuccess) return e;
        out->fe_mul_fixed_per_s = threads * kMulIters * 2.0 / (ms * 1e-3);
    }
    // bandwidth: 2 GiB in, 2 GiB out (>> 126 MB L2)
    const uint64_t n = (uint64_t)1 << 26;
    Fe *a = nullptr, *b = nullptr;
    e = cudaMalloc((void**)&a, n * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b, n * 32);
    if (e == cudaSuccess) e = cudaMemsetAsync(a, 1, n * 32, st);
    if (e == cudaSuccess) e = time_ms([&] { copy_kernel<<<sms * 16, kThreads, 0, st>>>(a, b, n); }, st, 5, &ms);
    if (e == cudaSuccess) out->copy_gbs = 2.0 * n * 32 / (ms * 1e-3) / 1e9;
    if (e == cudaSuccess) e = time_ms([&] { read_kernel<<<sms * 16, kThreads, 0, st>>>(a, n, (Fe*)sink); }, st, 5, &ms);
    if (e == cudaSuccess) out->read_gbs = 1.0 * n * 32 / (ms * 1e-3) / 1e9;
    cudaFree(a);
    cudaFree(b);
    cudaFree(sink);
    cudaFree(clk);
    return e;
}

}  // namespace zk
