// kernels_mle.cu — MLE utilities around the hot path (sm_100a): general-variable partial evaluation,
// element-wise product, canonical big-endian serialisation, Montgomery conversion, the synthetic
// table generator and the multi-GPU glue kernels.  Reference call sites cited per kernel.
#include "kernels.h"
#include "host_field.hpp"

namespace zk {
namespace {

constexpr int kThreads = 256;

inline unsigned grid_1d(uint64_t items, unsigned cap = 148 * 16) {
    uint64_t need = (items + kThreads - 1) / kThreads;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// MultiLinearPolynomial::partial_evaluate, one assignment, any variable
// (polynomial/src/multilinear/evaluation_form.rs:54-72; pair addressing pairing_index.rs:2-21):
// out[k] = L - a (L - R), L = in[insert_bit(k,pos,0)], R = in[L_index | 1<<pos].
template <class F>
__global__ void __launch_bounds__(kThreads)
    fold_var_kernel(const Fe* in, Fe* out, uint64_t pairs, unsigned pos, const __grid_constant__ FixedMul atab) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads, low_mask = ((uint64_t)1 << pos) - 1;
    for (uint64_t k = (uint64_t)blockIdx.x * kThreads + threadIdx.x; k < pairs; k += stride) {
        uint64_t l = ((k >> pos) << (pos + 1)) | (k & low_mask);
        Fe L = ld_fe(in + l), R = ld_fe(in + (l | ((uint64_t)1 << pos)));
        st_fe(out + k, fe_fold_fixed<F>(L, R, atab));
    }
}

// ProductPoly::prod_reduce (polynomial/src/product_poly.rs:66-74)
template <class F>
__global__ void __launch_bounds__(kThreads) prod_reduce_kernel(TablePtrs tabs, int m, uint64_t n, Fe* out) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
        Fe pr = ld_fe(tabs.t[0] + j);
#pragma unroll 1
        for (int k = 1; k < m; k++) pr = fe_mul<F>(pr, ld_fe(tabs.t[k] + j));
        st_fe(out + j, pr);
    }
}

// MultiLinearPolynomial::to_bytes (evaluation_form.rs:97-103): into_bigint().to_bytes_be(), 32 B per element
template <class F>
__global__ void __launch_bounds__(kThreads) to_bytes_kernel(const Fe* in, uint64_t n, Fe* out) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
        Fe c = fe_to_canonical<F>(ld_fe(in + j));
        Fe be;
#pragma unroll
        for (int i = 0; i < 8; i++) be.v[i] = __byte_perm(c.v[7 - i], 0, 0x0123);
        st_fe(out + j, be);
    }
}

template <class F>
__global__ void __launch_bounds__(kThreads) convert_kernel(Fe* data, uint64_t n, bool to_mont) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += stride) {
        Fe x = ld_fe(data + j);
        st_fe(data + j, to_mont ? fe_from_canonical<F>(x) : fe_to_canonical<F>(x));
    }
}

__device__ __forceinline__ uint64_t splitmix64_mix(uint64_t z) {
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
// Synthetic tables: canonical v < 2^254 from a counter-based splitmix64 keyed by the
// GLOBAL index, stored in Montgomery form.
template <class F>
__global__ void __launch_bounds__(kThreads)
    generate_kernel(Fe* out, uint64_t count, uint64_t seed, uint64_t table_id, uint64_t first, uint64_t gstride) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t j = (uint64_t)blockIdx.x * kThreads + threadIdx.x; j < count; j += stride) {
        const uint64_t i = first + j * gstride;
        Fe c;
#pragma unroll
        for (int l = 0; l < 4; l++) {
            uint64_t ctr = ((table_id << 40) + i) * 4 + (uint64_t)l;
            uint64_t w = splitmix64_mix(seed + 0x9E3779B97F4A7C15ULL * (ctr + 1));
            if (l == 3) w &= 0x3FFFFFFFFFFFFFFFULL;
            c.v[2 * l] = (uint32_t)w;
            c.v[2 * l + 1] = (uint32_t)(w >> 32);
        }
        if (F::ID == Fr377::ID) {  // 2^254 < 5p for the 253-bit modulus: bring v into [0,p)
#pragma unroll 1
            for (int it = 0; it < 4; it++) c = fe_reduce_once<F>(c);
        }
        st_fe(out + j, fe_from_canonical<F>(c));
    }
}

// `batch` tables back to back (in: [b][q][j], out: [b][j * world + q]) in one launch
__global__ void __launch_bounds__(kThreads) interleave_kernel(const Fe* in, Fe* out, uint64_t local_len, unsigned world, unsigned batch) {
    const uint64_t per = local_len * world, total = per * batch, stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t g = (uint64_t)blockIdx.x * kThreads + threadIdx.x; g < total; g += stride) {
        const uint64_t b = g / per, r = g % per, j = r / world, q = r % world;
        st_fe(out + g, ld_fe(in + b * per + q * local_len + j));
    }
}

// lanes hold sums of <= 2^24 32-bit limbs: carry-propagate to a (<= 280-bit) integer, reduce mod p.
template <class F>
__device__ __forceinline__ void narrow_one(const uint64_t* lanes, Fe* out_dev, Fe* out_host, int e) {
    uint32_t w[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += lanes[e * 8 + i];
        w[i] = (uint32_t)c;
        c >>= 32;
    }
    w[8] = (uint32_t)c;  // value = w[8]*2^256 + w[0..7], w[8] < world size
    // subtract p while value >= p (value < world * p, world <= 2^16 in principle; loop is short)
    while (true) {
        uint32_t t[8];
        int64_t b = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int64_t d = (int64_t)w[i] - (int64_t)F::p(i) + b;
            t[i] = (uint32_t)d;
            b = d >> 32;
        }
        int64_t top = (int64_t)w[8] + b;
        if (top < 0) break;
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = t[i];
        w[8] = (uint32_t)top;
    }
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    out_dev[e] = r;
    out_host[e] = r;
}

template <class F>
__global__ void narrow_kernel(const uint64_t* lanes, Fe* out_dev, Fe* out_host, int count, unsigned* flag_host,
                              unsigned seq) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < count) narrow_one<F>(lanes, out_dev, out_host, e);
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0 && seq != 0) {
        *(volatile unsigned*)flag_host = seq;
        __threadfence_system();
    }
}
}  // namespace

#define ZK_FIELD_DISPATCH(field, CALL381, CALL377) \
    do {                                           \
        if ((field) == Fr381::ID) { CALL381; }     \
        else { CALL377; }                          \
    } while (0)

cudaError_t launch_fold_var(int field, const Fe* in, Fe* out, unsigned nv, unsigned initial_var, const Fe& a,
                            cudaStream_t stream, int* launches) {
    const unsigned pos = nv - 1 - initial_var;
    const uint64_t pairs = (uint64_t)1 << (nv - 1);
    host::Field HF(field);
    host::El am;
    std::memcpy(am.v, a.v, 32);
    FixedMul atab;
    host::fixed_mul_table(HF, am, atab.v);  // the assignment is the same for every pair: fe_mul_fixed
    ZK_FIELD_DISPATCH(field, (fold_var_kernel<Fr381><<<grid_1d(pairs), kThreads, 0, stream>>>(in, out, pairs, pos, atab)),
                      (fold_var_kernel<Fr377><<<grid_1d(pairs), kThreads, 0, stream>>>(in, out, pairs, pos, atab)));
    ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_prod_reduce(int field, const TablePtrs& tabs, int m, uint64_t n, Fe* out, cudaStream_t stream,
                               int* launches) {
    ZK_FIELD_DISPATCH(field, (prod_reduce_kernel<Fr381><<<grid_1d(n), kThreads, 0, stream>>>(tabs, m, n, out)),
                      (prod_reduce_kernel<Fr377><<<grid_1d(n), kThreads, 0, stream>>>(tabs, m, n, out)));
    ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_to_bytes(int field, const Fe* in, uint64_t n, uint8_t* out, cudaStream_t stream, int* launches) {
    ZK_FIELD_DISPATCH(field, (to_bytes_kernel<Fr381><<<grid_1d(n), kThreads, 0, stream>>>(in, n, (Fe*)out)),
                      (to_bytes_kernel<Fr377><<<grid_1d(n), kThreads, 0, stream>>>(in, n, (Fe*)out)));
    ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_convert(int field, Fe* data, uint64_t n, bool to_mont, cudaStream_t stream, int* launches) {
    ZK_FIELD_DISPATCH(field, (convert_kernel<Fr381><<<grid_1d(n), kThreads, 0, stream>>>(data, n, to_mont)),
                      (convert_kernel<Fr377><<<grid_1d(n), kThreads, 0, stream>>>(data, n, to_mont)));
    ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_generate(int field, Fe* out, uint64_t count, uint64_t seed, uint64_t table_id, uint64_t first,
                            uint64_t stride, cudaStream_t stream, int* launches) {
    ZK_FIELD_DISPATCH(
        field, (generate_kernel<Fr381><<<grid_1d(count), kThreads, 0, stream>>>(out, count, seed, table_id, first, stride)),
        (generate_kernel<Fr377><<<grid_1d(count), kThreads, 0, stream>>>(out, count, seed, table_id, first, stride)));
    ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_interleave(const Fe* in, Fe* out, uint64_t local_len, unsigned world, cudaStream_t stream,
                              int* launches, unsigned batch) {
    interleave_kernel<<<grid_1d(local_len * world * batch), kThreads, 0, stream>>>(in, out, local_len, world, batch);
    ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_narrow(int field, const uint64_t* lanes, Fe* out_dev, Fe* out_host_devptr, int count,
                          unsigned* flag_host_devptr, unsigned seq, cudaStream_t stream, int* launches) {
    ZK_FIELD_DISPATCH(field, (narrow_kernel<Fr381><<<1, 32, 0, stream>>>(lanes, out_dev, out_host_devptr, count, flag_host_devptr, seq)),
                      (narrow_kernel<Fr377><<<1, 32, 0, stream>>>(lanes, out_dev, out_host_devptr, count, flag_host_devptr, seq)));
    ++*launches;
    return cudaGetLastError();
}

}  // namespace zk
