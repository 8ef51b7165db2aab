// ntt_sharded_kernels.cuh — the three small kernels the multi-GPU NTT adds around the single-GPU transform
// (kernels_ntt.cu): a power table, the inter-rank twiddle multiplication, and the G-point DFT across ranks.
//
// Factorisation (G = 2^g ranks, N = 2^n points, M = N / G, rank q holds the strided shard a_q[j] = a[j G + q]):
//     X[k' + M c] = sum_q w_G^(q c) * [ w_N^(q k') * A_q[k'] ],      A_q = NTT_M(a_q),   k' < M, c < G
// i.e. fft/src/lib.rs:21-46 unrolled g levels from the top: a local M-point transform per rank, one twiddle
// multiplication per element, and a G-point DFT over the ranks' values — which an all-to-all turns into local work.
// Kept in a header so that tests/cpp/test_ntt_sharded_host.cpp can replay the very same source on the host.
// Needs in scope: Fe and the fe_* / ld_fe / st_fe functions (field.cuh on the device).
#pragma once

namespace zk {
namespace {

constexpr int kShThreads = 256;
constexpr int kMaxRanks = 8;

// out[i] = base^(i << shift), i < count  (square-and-multiply per entry: the tables are a few thousand entries)
template <class F>
__global__ void __launch_bounds__(kShThreads) pow_table_kernel(Fe* out, uint64_t count, Fe base, unsigned shift) {
    const uint64_t stride = (uint64_t)gridDim.x * kShThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kShThreads + threadIdx.x; i < count; i += stride) {
        uint64_t e = i << shift;
        Fe acc = fe_one<F>(), b = base;
        while (e) {
            if (e & 1) acc = fe_mul<F>(acc, b);
            e >>= 1;
            if (e) b = fe_mul<F>(b, b);
        }
        st_fe(out + i, acc);
    }
}

// x[k] *= w^k with w^k = t_hi[k >> lo_bits] * t_lo[k & (2^lo_bits - 1)]   (t_lo[i] = w^i, t_hi[i] = w^(i << lo_bits))
template <class F>
__global__ void __launch_bounds__(kShThreads)
    twiddle_mul_kernel(Fe* x, uint64_t m, const Fe* t_lo, const Fe* t_hi, unsigned lo_bits) {
    const uint64_t stride = (uint64_t)gridDim.x * kShThreads, mask = ((uint64_t)1 << lo_bits) - 1;
    for (uint64_t k = (uint64_t)blockIdx.x * kShThreads + threadIdx.x; k < m; k += stride) {
        const Fe w = fe_mul<F>(ld_fe(t_hi + (k >> lo_bits)), ld_fe(t_lo + (k & mask)));
        st_fe(x + k, fe_mul<F>(ld_fe(x + k), w));
    }
}

struct GdftParams {
    Fe w[kMaxRanks / 2];  // w_G^(+-i), i < G/2 (Montgomery form)
    Fe scale;             // G^-1 (inverse transform)
    int do_scale;
};

// out[c][j] = scale * sum_q w^(q c) in[q][j]  for j < chunk: a radix-2 DIF on G values held in registers; its
// outputs come out bit-reversed and are stored to their natural slot.
template <class F, int G>
__global__ void __launch_bounds__(kShThreads)
    gdft_kernel(const Fe* in, Fe* out, uint64_t chunk, const __grid_constant__ GdftParams prm) {
    static_assert(G == 2 || G == 4 || G == 8, "ranks: 2, 4 or 8");
    const uint64_t stride = (uint64_t)gridDim.x * kShThreads;
    for (uint64_t j = (uint64_t)blockIdx.x * kShThreads + threadIdx.x; j < chunk; j += stride) {
        Fe y[G];
#pragma unroll
        for (int q = 0; q < G; q++) y[q] = ld_fe(in + (uint64_t)q * chunk + j);
#pragma unroll
        for (int h = G / 2; h >= 1; h /= 2) {
#pragma unroll
            for (int b = 0; b < G; b += 2 * h) {
#pragma unroll
                for (int i = 0; i < h; i++) {
                    const Fe u = y[b + i], v = y[b + i + h];
                    y[b + i] = fe_add<F>(u, v);
                    const Fe d = fe_sub<F>(u, v);
                    const int tw = i * (G / (2 * h));  // exponent of w_G, < G/2
                    y[b + i + h] = tw == 0 ? d : fe_mul<F>(d, prm.w[tw]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < G; c++) {
            int rev = 0;
#pragma unroll
            for (int bit = 1, rb = G / 2; bit < G; bit <<= 1, rb >>= 1)
                if (c & bit) rev |= rb;
            Fe v = y[rev];
            if (prm.do_scale) v = fe_mul<F>(v, prm.scale);
            st_fe(out + (uint64_t)c * chunk + j, v);
        }
    }
}

}  // namespace
}  // namespace zk
