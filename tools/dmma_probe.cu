// dmma_probe.cu — does the FP64 tensor path (mma.sync.m8n8k4.f64 -> DMMA.8x8x4 on sm_100a) sustain at least the vector DFMA
// rate on this GPU, and is it exact on integer operands below 2^53?  (The fold's column sums are such integers.)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dmma_probe tools/dmma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) dmma_kernel(double* sink, int ilp_sel) {
    double a = 1.0 + (threadIdx.x & 3), b = 2.0 + (threadIdx.x >> 2);
    double c[8][2];
    for (int i = 0; i < 8; i++) { c[i][0] = 4503599627370496.0; c[i][1] = 4503599627370496.0 + i; }
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 0.123) sink[0] = s;
}
__global__ void __launch_bounds__(256) dfma_kernel(double* sink) {
    double a = 1.0 + (threadIdx.x & 3), b = 1e-9 * threadIdx.x;
    double c[16];
    for (int i = 0; i < 16; i++) c[i] = i;
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 0.123) sink[0] = s;
}
// exactness: C = A(8x4) * B(4x8) with 16-bit A and 32-bit B entries, accumulated 4 times onto 2^52
__global__ void exact_kernel(const uint32_t* a16, const uint32_t* b32, unsigned long long* out) {
    const int lane = threadIdx.x;
    double c0 = 4503599627370496.0, c1 = c0;
    for (int ks = 0; ks < 4; ks++) {
        const double A = (double)a16[ks * 32 + lane];   // A[row = lane>>2][col = lane&3]
        const double B = (double)b32[ks * 32 + lane];   // B[row = lane&3][col = lane>>2]
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(A), "d"(B));
    }
    out[2 * lane] = (unsigned long long)__double_as_longlong(c0) & 0xfffffffffffffull;
    out[2 * lane + 1] = (unsigned long long)__double_as_longlong(c1) & 0xfffffffffffffull;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink;
    cudaMalloc(&sink, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8;
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); dmma_kernel<<<grid, 256>>>(sink, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double dmma = (double)grid * 8 /*warps*/ * kIters * 8 / (ms * 1e-3);
    printf("DMMA.8x8x4: %.3e warp-instr/s = %.3e FMA/s  (%.2f ms)\n", dmma, dmma * 256, ms);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); dfma_kernel<<<grid, 256>>>(sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double dfma = (double)grid * 256 * kIters * 16 / (ms * 1e-3);
    printf("DFMA      : %.3e FMA/s  (%.2f ms)\n", dfma, ms);
    // exactness
    uint32_t ha[128], hb[128];
    unsigned long long hout[64], want[64];
    uint64_t st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    for (int i = 0; i < 128; i++) { ha[i] = (i % 7 == 0) ? 0xffffu : (uint32_t)(rnd() & 0xffff); hb[i] = (i % 5 == 0) ? 0xffffffffu : (uint32_t)rnd(); }
    for (int row = 0; row < 8; row++)
        for (int col = 0; col < 8; col++) {
            unsigned long long acc = 0;
            for (int ks = 0; ks < 4; ks++)
                for (int k = 0; k < 4; k++) acc += (unsigned long long)ha[ks * 32 + row * 4 + k] * hb[ks * 32 + col * 4 + k];
            want[row * 8 + col] = acc;
        }
    uint32_t *da, *db; unsigned long long* dout;
    cudaMalloc(&da, 512); cudaMalloc(&db, 512); cudaMalloc(&dout, 512);
    cudaMemcpy(da, ha, 512, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, 512, cudaMemcpyHostToDevice);
    exact_kernel<<<1, 32>>>(da, db, dout);
    cudaMemcpy(hout, dout, 512, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int lane = 0; lane < 32; lane++)
        for (int j = 0; j < 2; j++) bad += hout[2 * lane + j] != want[(lane >> 2) * 8 + (lane & 3) * 2 + j];
    printf("exactness: %d mismatches of 64 (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
    return bad != 0;
}
