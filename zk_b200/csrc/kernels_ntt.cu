// kernels_ntt.cu — radix-2 NTT / INTT over the 256-bit fields (secondary hot path).
//
// Replaces fft/src/lib.rs:4-61 (`fft`, `ifft`, `fft_internal`, `split_even_odd`): X[i] = sum_j a_j w^(ij),
// natural order in, natural order out, w = g^((p-1)/N) (w^-1 and a final N^-1 scale for the inverse).
// The reference recurses with fresh vectors, recomputes `omega.pow([i])` twice per butterfly and
// inverts N once per output element; here twiddles w^i (i < N/2) are built once per plan by doubling,
// the transform is log2(N) decimation-in-frequency stages (one twiddle multiplication per butterfly,
// X - Y*w reuses the product) followed by one bit-reversal pass that also applies N^-1.
// Any exact algorithm gives bit-identical outputs (integer arithmetic).
#include "keccak.hpp"  // host_field.hpp
#include "kernels.h"

namespace zk {

struct NttPlan {
    int field;
    unsigned log_n;
    bool inverse;
    Fe* twiddles;  // w^i, i < N/2 (Montgomery)
    Fe n_inv;      // N^-1 (inverse only)
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

template <class F>
cudaError_t plan_build(NttPlan* p, cudaStream_t st, int* launches) {
    host::Field HF(F::ID);
    host::El w = HF.root_of_unity(p->log_n);
    if (p->inverse) {
        w = HF.inverse(w);
        host::El ninv = HF.inverse(HF.from_u64((uint64_t)1 << p->log_n));
        std::memcpy(p->n_inv.v, ninv.v, 32);
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
cudaError_t execute(NttPlan* p, Fe* data, cudaStream_t st, int* launches) {
    const uint64_t n = (uint64_t)1 << p->log_n, n_half = n >> 1;
    for (unsigned s = 0; s < p->log_n; s++) {
        dif_stage_kernel<F><<<grid_1d(n_half), kThreads, 0, st>>>(data, p->twiddles, n_half, p->log_n - 1 - s, s);
        ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    bitrev_kernel<F><<<grid_1d(n), kThreads, 0, st>>>(data, n, p->log_n, p->inverse, p->n_inv);
    ++*launches;
    return cudaGetLastError();
}
}  // namespace

cudaError_t ntt_plan_create(int field, unsigned log_n, bool inverse, cudaStream_t stream, NttPlan** out, int* launches) {
    NttPlan* p = new NttPlan{field, log_n, inverse, nullptr, Fe{}};
    cudaError_t e = field == Fr381::ID ? plan_build<Fr381>(p, stream, launches) : plan_build<Fr377>(p, stream, launches);
    if (e != cudaSuccess) {
        ntt_plan_destroy(p);
        return e;
    }
    *out = p;
    return cudaSuccess;
}
void ntt_plan_destroy(NttPlan* p) {
    if (!p) return;
    cudaFree(p->twiddles);
    delete p;
}
cudaError_t ntt_execute(NttPlan* plan, Fe* data, cudaStream_t stream, int* launches) {
    return plan->field == Fr381::ID ? execute<Fr381>(plan, data, stream, launches)
                                    : execute<Fr377>(plan, data, stream, launches);
}

}  // namespace zk
