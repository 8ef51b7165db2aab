// kernels_ntt_sharded.cu — launchers of the multi-GPU NTT's own kernels (ntt_sharded_kernels.cuh); the transform of
// each rank's shard is the single-GPU NTT of kernels_ntt.cu, the exchanges are NCCL send/recv groups (api_ntt.cu).
#include "kernels.h"
#include "ntt_sharded_kernels.cuh"

namespace zk {
namespace {

inline unsigned sh_grid(uint64_t items, unsigned cap = 148 * 8) {
    uint64_t need = (items + kShThreads - 1) / kShThreads;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// out[q * L + j] = in[j * G + q]: the strided shards of a full table, rank-major (inverse of launch_interleave)
__global__ void __launch_bounds__(kShThreads) deinterleave_kernel(const Fe* in, Fe* out, uint64_t local_len, unsigned world) {
    const uint64_t total = local_len * world, stride = (uint64_t)gridDim.x * kShThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kShThreads + threadIdx.x; i < total; i += stride) {
        const uint64_t q = i / local_len, j = i - q * local_len;
        st_fe(out + i, ld_fe(in + j * world + q));
    }
}

template <class F>
cudaError_t gdft_dispatch(int ranks, const Fe* in, Fe* out, uint64_t chunk, const GdftParams& prm, cudaStream_t st) {
    const unsigned grid = sh_grid(chunk);
    switch (ranks) {
        case 2: gdft_kernel<F, 2><<<grid, kShThreads, 0, st>>>(in, out, chunk, prm); break;
        case 4: gdft_kernel<F, 4><<<grid, kShThreads, 0, st>>>(in, out, chunk, prm); break;
        case 8: gdft_kernel<F, 8><<<grid, kShThreads, 0, st>>>(in, out, chunk, prm); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_pow_table(int field, Fe* out, uint64_t count, const Fe& base, unsigned shift, cudaStream_t stream,
                             int* launches) {
    ++*launches;
    if (field == Fr381::ID) pow_table_kernel<Fr381><<<sh_grid(count), kShThreads, 0, stream>>>(out, count, base, shift);
    else pow_table_kernel<Fr377><<<sh_grid(count), kShThreads, 0, stream>>>(out, count, base, shift);
    return cudaGetLastError();
}

cudaError_t launch_twiddle_mul(int field, Fe* x, uint64_t m, const Fe* t_lo, const Fe* t_hi, unsigned lo_bits,
                               cudaStream_t stream, int* launches) {
    ++*launches;
    if (field == Fr381::ID) twiddle_mul_kernel<Fr381><<<sh_grid(m), kShThreads, 0, stream>>>(x, m, t_lo, t_hi, lo_bits);
    else twiddle_mul_kernel<Fr377><<<sh_grid(m), kShThreads, 0, stream>>>(x, m, t_lo, t_hi, lo_bits);
    return cudaGetLastError();
}

cudaError_t launch_gdft(int field, int ranks, const Fe* in, Fe* out, uint64_t chunk, const Fe* w_half, const Fe* scale,
                        cudaStream_t stream, int* launches) {
    if (ranks != 2 && ranks != 4 && ranks != 8) return cudaErrorInvalidValue;
    GdftParams prm{};
    for (int i = 0; i < ranks / 2; i++) prm.w[i] = w_half[i];
    prm.do_scale = scale ? 1 : 0;
    if (scale) prm.scale = *scale;
    ++*launches;
    return field == Fr381::ID ? gdft_dispatch<Fr381>(ranks, in, out, chunk, prm, stream)
                              : gdft_dispatch<Fr377>(ranks, in, out, chunk, prm, stream);
}

cudaError_t launch_deinterleave(const Fe* in, Fe* out, uint64_t local_len, unsigned world, cudaStream_t stream,
                                int* launches) {
    ++*launches;
    deinterleave_kernel<<<sh_grid(local_len * world), kShThreads, 0, stream>>>(in, out, local_len, world);
    return cudaGetLastError();
}

}  // namespace zk
