// kernels_ntt.cu — radix-2 NTT / INTT over the 256-bit fields (secondary hot path).
//
// Replaces fft/src/lib.rs:4-61 (`fft`, `ifft`, `fft_internal`, `split_even_odd`): X[i] = sum_j a_j w^(ij),
// natural order in, natural order out, w = g^((p-1)/N) (w^-1 and a final N^-1 scale for the inverse).
// The reference recurses with fresh vectors, recomputes `omega.pow([i])` twice per butterfly and
// inverts N once per output element; here twiddles w^i (i < N/2) are built once per plan by doubling,
// the transform is log2(N) decimation-in-frequency stages (one twiddle multiplication per butterfly,
// X - Y*w reuses the product) followed by one bit-reversal pass that also applies N^-1.  Stages are grouped
// into passes of up to 9: a pass loads tiles of 2^s elements (stride N_b/2^s, 4 adjacent tiles per block so
// global accesses are 128-byte segments) into shared memory, runs a plain 2^s-point DIF there with the small
// twiddles w_{2^s}^j, multiplies output c by the inter-pass twiddle w_N^(2^t0 * base * c) and writes back in
// place (the classic four-step factorisation).  The last pass writes each element directly to its natural-order
// (bit-reversed) slot of a second buffer, scaled by N^-1 for the inverse, and the caller swaps the two buffers:
// 2^28 points take 4 passes over the data instead of 28 + a permutation pass.
// Any exact algorithm gives bit-identical outputs (integer arithmetic).
#include "keccak.hpp"  // host_field.hpp
#include "kernels.h"

#include <cstdlib>

namespace zk {

struct NttPlan {
    int field;
    unsigned log_n;
    bool inverse;
    Fe* twiddles;  // w^i, i < N/2 (Montgomery)
    Fe n_inv;      // N^-1 (inverse only)
    Fe* scratch;   // N elements: the last pass writes the natural-order result here (lazily allocated)
    FixedMul n_inv_tab;  // multiples of N^-1 for fe_mul_fixed (inverse only)
};

namespace {
constexpr int kThreads = 256;
inline unsigned grid_1d(uint64_t items, unsigned cap = 148 * 16) {
    uint64_t need = (items + kThreads - 1) / kThreads;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// tw[count + i] = tw[i] * w_pow   (w_pow = w^count): doubles the table
template <class F>
__global__ void __launch_bounds__(kThreads) twiddle_extend_kernel(Fe* tw, uint64_t count, Fe w_pow) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += stride)
        st_fe(tw + count + i, fe_mul<F>(ld_fe(tw + i), w_pow));
}

// One DIF stage: butterflies (i0, i0+half): a[i0] = u+v, a[i1] = (u-v) * w^(j << s)
template <class F>
__global__ void __launch_bounds__(kThreads)
    dif_stage_kernel(Fe* a, const Fe* __restrict__ tw, uint64_t n_half, unsigned log_half, unsigned s) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads, half = (uint64_t)1 << log_half;
    for (uint64_t k = (uint64_t)blockIdx.x * kThreads + threadIdx.x; k < n_half; k += stride) {
        const uint64_t j = k & (half - 1), i0 = ((k >> log_half) << (log_half + 1)) | j, i1 = i0 + half;
        Fe u = ld_fe(a + i0), v = ld_fe(a + i1);
        st_fe(a + i0, fe_add<F>(u, v));
        Fe d = fe_sub<F>(u, v);
        st_fe(a + i1, j == 0 ? d : fe_mul<F>(d, ld_fe(tw + (j << s))));
    }
}

// In-place bit reversal (+ optional scale by n_inv)
template <class F>
__global__ void __launch_bounds__(kThreads) bitrev_kernel(Fe* a, uint64_t n, unsigned log_n, bool scale, Fe n_inv) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        const uint64_t j = __brevll(i) >> (64 - log_n);
        if (i < j) {
            Fe x = ld_fe(a + i), y = ld_fe(a + j);
            if (scale) { x = fe_mul<F>(x, n_inv); y = fe_mul<F>(y, n_inv); }
            st_fe(a + i, y);
            st_fe(a + j, x);
        } else if (i == j && scale) {
            st_fe(a + i, fe_mul<F>(ld_fe(a + i), n_inv));
        }
    }
}


// ---- shared-memory pass: s consecutive DIF stages starting at stage t0 ------------------------------
constexpr int kTileB = 4;          // adjacent tiles per block (128-byte global segments)
constexpr int kMaxTileLog = 9;     // 2^9 * 4 * 32 B = 64 KB of shared memory per block
constexpr int kPassThreads = 256;

__device__ __forceinline__ Fe lds_fe(const uint4* lo, const uint4* hi, unsigned i) {
    uint4 a = lo[i], b = hi[i];
    Fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fe(uint4* lo, uint4* hi, unsigned i, const Fe& r) {
    lo[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

template <class F>
__global__ void __launch_bounds__(kPassThreads)
    ntt_pass_kernel(Fe* a, const Fe* __restrict__ tw, unsigned log_n, unsigned t0, unsigned s, uint64_t n_groups,
                    Fe* out_natural, bool scale, const __grid_constant__ FixedMul n_inv_tab) {
    extern __shared__ uint4 smem[];
    const unsigned tile = 1u << s, elems = tile * kTileB;
    uint4* lo = smem;                 // [elems]   limbs 0-3 of element (x, b) at x*kTileB + b
    uint4* hi = smem + elems;         // [elems]   limbs 4-7
    uint4* wlo = smem + 2 * elems;    // [tile/2]  small twiddles w_{2^s}^j
    uint4* whi = wlo + (tile >> 1);
    const unsigned log_nb = log_n - t0;                  // this pass works inside blocks of 2^log_nb
    const unsigned log_stride = log_nb - s;              // distance between consecutive tile elements
    const uint64_t stride = (uint64_t)1 << log_stride;
    const bool b_fastest = stride >= kTileB;             // which index walks contiguous memory
    for (unsigned j = threadIdx.x; j < (tile >> 1); j += kPassThreads)
        sts_fe(wlo, whi, j, ld_fe(tw + ((uint64_t)j << (log_n - s))));
    for (uint64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const uint64_t tile0 = g * kTileB;               // first of kTileB consecutive tile ids
        __syncthreads();
        // ---- load ----
        for (unsigned e = threadIdx.x; e < elems; e += kPassThreads) {
            unsigned x, b;
            if (b_fastest) { b = e % kTileB; x = e / kTileB; } else { x = e % tile; b = e / tile; }
            const uint64_t t = tile0 + b, blk = t >> log_stride, base = t & (stride - 1);
            sts_fe(lo, hi, x * kTileB + b, ld_fe_stream(a + (blk << log_nb) + base + ((uint64_t)x << log_stride)));
        }
        __syncthreads();
        // ---- s DIF stages in shared memory ----
        for (unsigned u = 0; u < s; u++) {
            const unsigned log_half = s - 1 - u, half = 1u << log_half;
            for (unsigned q = threadIdx.x; q < (elems >> 1); q += kPassThreads) {
                const unsigned b = q % kTileB, j = q / kTileB, jl = j & (half - 1);
                const unsigned x0 = ((j >> log_half) << (log_half + 1)) | jl, x1 = x0 + half;
                Fe p = lds_fe(lo, hi, x0 * kTileB + b), v = lds_fe(lo, hi, x1 * kTileB + b);
                sts_fe(lo, hi, x0 * kTileB + b, fe_add<F>(p, v));
                Fe d = fe_sub<F>(p, v);
                if (jl != 0) d = fe_mul<F>(d, lds_fe(wlo, whi, jl << u));
                sts_fe(lo, hi, x1 * kTileB + b, d);
            }
            __syncthreads();
        }
        // ---- inter-pass twiddle + store (in place) ----
        for (unsigned e = threadIdx.x; e < elems; e += kPassThreads) {
            unsigned x, b;
            if (b_fastest) { b = e % kTileB; x = e / kTileB; } else { x = e % tile; b = e / tile; }
            const uint64_t t = tile0 + b, blk = t >> log_stride, base = t & (stride - 1);
            Fe val = lds_fe(lo, hi, x * kTileB + b);
            if (log_stride != 0) {
                const uint64_t c = __brev(x) >> (32 - s);                  // position x holds output c = bitrev_s(x)
                const uint64_t ex = (base * c) << t0, half_n = (uint64_t)1 << (log_n - 1);
                if (ex != 0) {
                    if (ex >= half_n) val = fe_sub<F>(fe_zero<F>(), fe_mul<F>(val, ld_fe(tw + (ex - half_n))));
                    else val = fe_mul<F>(val, ld_fe(tw + ex));
                }
            }
            const uint64_t g = (blk << log_nb) + base + ((uint64_t)x << log_stride);
            if (out_natural != nullptr) {
                // last pass: the element at position g of the in-place DIF is X[bitrev(g)] — write it straight to
                // its natural-order slot (32-byte sector writes) and fold the N^-1 of the inverse transform in
                if (scale) val = fe_mul_fixed<F>(val, n_inv_tab);
                st_fe(out_natural + (__brevll(g) >> (64 - log_n)), val);
            } else {
                st_fe(a + g, val);
            }
        }
    }
}

template <class F>
cudaError_t plan_build(NttPlan* p, cudaStream_t st, int* launches) {
    host::Field HF(F::ID);
    host::El w = HF.root_of_unity(p->log_n);
    if (p->inverse) {
        w = HF.inverse(w);
        host::El ninv = HF.inverse(HF.from_u64((uint64_t)1 << p->log_n));
        std::memcpy(p->n_inv.v, ninv.v, 32);
        host::fixed_mul_table(HF, ninv, p->n_inv_tab.v);
    }
    const uint64_t half_n = (uint64_t)1 << (p->log_n - 1);
    cudaError_t e = cudaMalloc((void**)&p->twiddles, (size_t)half_n * sizeof(Fe));
    if (e != cudaSuccess) return e;
    host::El one = HF.one();
    e = cudaMemcpyAsync(p->twiddles, one.v, 32, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    host::El wp = w;  // w^(2^b)
    for (uint64_t count = 1; count < half_n; count <<= 1) {
        Fe wf;
        std::memcpy(wf.v, wp.v, 32);
        twiddle_extend_kernel<F><<<grid_1d(count), kThreads, 0, st>>>(p->twiddles, count, wf);
        ++*launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        wp = HF.mul(wp, wp);
    }
    return cudaStreamSynchronize(st);
}

template <class F>
cudaError_t execute(NttPlan* p, Fe* data, Fe** result, cudaStream_t st, int* launches) {
    *result = data;
    const unsigned k = p->log_n;
    const uint64_t n = (uint64_t)1 << k;
    // split the k stages into ceil(k/9) passes of (almost) equal size, every pass >= 2 stages when k >= 4
    const unsigned n_pass = (k + kMaxTileLog - 1) / kMaxTileLog;
    const size_t max_smem = (size_t)(2 * ((1u << kMaxTileLog) * kTileB) + (1u << kMaxTileLog)) * sizeof(uint4);
    static PerDeviceCache attr_cache;  // the shared-memory opt-in is per device
    const int attr = per_device(attr_cache, [max_smem] {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
        return e == cudaSuccess ? 1 : -(int)e;
    });
    if (attr <= 0) return (cudaError_t)(-attr);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned t0 = 0;
    for (unsigned pi = 0; pi < n_pass; pi++) {
        const unsigned s = (k - t0) / (n_pass - pi);       // remaining stages spread over remaining passes
        const uint64_t n_tiles = n >> s;
        if (n_tiles < kTileB) {                            // tiny transform: the per-stage kernel is enough
            for (unsigned t = t0; t < t0 + s; t++) {
                dif_stage_kernel<F><<<grid_1d(n >> 1), kThreads, 0, st>>>(data, p->twiddles, n >> 1, k - 1 - t, t);
                ++*launches;
            }
        } else {
            const uint64_t n_groups = n_tiles / kTileB;
            const size_t smem = (size_t)(2 * ((1u << s) * kTileB) + (1u << s)) * sizeof(uint4);
            // One block per group of tiles (up to 1024 per SM) rather than a persistent grid of resident blocks: the
            // hardware block scheduler then balances the SMs (measured 4.08 ms against 4.66 ms at 2^24 with 3 per SM —
            // the same tail effect as in the round kernels).  ZK_B200_NTT_BPSM caps the blocks per SM for A/B runs.
            static const unsigned bpsm = [] {
                const char* e = std::getenv("ZK_B200_NTT_BPSM");
                const int v = e ? std::atoi(e) : 1024;
                return (unsigned)(v < 1 ? 1 : (v > 4096 ? 4096 : v));
            }();
            const unsigned grid = (unsigned)(n_groups < (uint64_t)sms * bpsm ? n_groups : (uint64_t)sms * bpsm);
            const bool last = (pi + 1 == n_pass);
            if (last) {
                if (!p->scratch) {
                    cudaError_t ea = cudaMalloc((void**)&p->scratch, (size_t)n * sizeof(Fe));
                    if (ea != cudaSuccess) return ea;
                }
                *result = p->scratch;
            }
            ntt_pass_kernel<F><<<grid, kPassThreads, smem, st>>>(data, p->twiddles, k, t0, s, n_groups,
                                                                last ? p->scratch : nullptr, p->inverse, p->n_inv_tab);
            ++*launches;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        t0 += s;
    }
    if (*result == data) {  // small transforms ran stage by stage in place: separate bit-reversal pass
        bitrev_kernel<F><<<grid_1d(n), kThreads, 0, st>>>(data, n, k, p->inverse, p->n_inv);
        ++*launches;
    }
    return cudaGetLastError();
}
}  // namespace

cudaError_t ntt_plan_create(int field, unsigned log_n, bool inverse, cudaStream_t stream, NttPlan** out, int* launches) {
    NttPlan* p = new NttPlan{field, log_n, inverse, nullptr, Fe{}, nullptr, FixedMul{}};
    cudaError_t e = field == Fr381::ID ? plan_build<Fr381>(p, stream, launches) : plan_build<Fr377>(p, stream, launches);
    if (e != cudaSuccess) {
        ntt_plan_destroy(p);
        return e;
    }
    *out = p;
    return cudaSuccess;
}
bool ntt_plan_is(const NttPlan* p, int field, unsigned log_n, bool inverse) {
    return p && p->field == field && p->log_n == log_n && p->inverse == inverse;
}
void ntt_plan_destroy(NttPlan* p) {
    if (!p) return;
    cudaFree(p->twiddles);
    cudaFree(p->scratch);
    delete p;
}
cudaError_t ntt_execute(NttPlan* plan, Fe* data, Fe** result, cudaStream_t stream, int* launches) {
    return plan->field == Fr381::ID ? execute<Fr381>(plan, data, result, stream, launches)
                                    : execute<Fr377>(plan, data, result, stream, launches);
}
void ntt_plan_adopt_scratch(NttPlan* plan, Fe* buf) { plan->scratch = buf; }

}  // namespace zk
