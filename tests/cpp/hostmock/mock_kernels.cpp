// TEST INFRASTRUCTURE ONLY — host stand-ins for the kernel launchers of zk_b200/csrc/kernels.h, used with
// mock_cudart.cpp to run api.cu's orchestration without a GPU.  Three kinds:
//   * the kernels that ARE replayable thread by thread run from their real source (sop_kernel.cuh,
//     ntt_sharded_kernels.cuh) with host_field.hpp as the arithmetic;
//   * the others (product round kernels, folds, generator, single-GPU NTT: shared memory, shuffles, PTX) are
//     replaced by naive models of their documented contract (kernels.h) — what is exercised there is the caller,
//     not the kernel: the GPU suite checks the kernels;
//   * the narrowing launch after the sharded all-reduce is modelled too (mock_nccl.cpp provides the collectives).
// Reducing launches publish like reduce_publish does: result_dev, the mapped host copy, the completion flag.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define __launch_bounds__(...)
// thread_local: the multi-rank driver runs one rank per thread
static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 gridDim, blockDim;

#include "kernels.h"
#include "field_f64.cuh"
#include "host_field.hpp"

namespace zk {
namespace {
thread_local const host::Field* g_field = nullptr;
thread_local host::El g_challenge;
thread_local std::vector<host::El> g_sums;
alignas(32) thread_local uint4 sop_smem[2 * 2 * (kMaxFactors + kMaxVirtual) * 128 + 5 * (4 * 128 + 128 / 4)];  // e/d arrays + wide accumulators (D <= 4)
constexpr int kThreads = 128;

inline host::El el(const Fe& a) { host::El e; std::memcpy(e.v, a.v, 32); return e; }
inline Fe fe(const host::El& e) { Fe a; std::memcpy(a.v, e.v, 32); return a; }
template <class F> Fe fe_zero() { return fe(g_field->zero()); }
template <class F> Fe fe_one() { return fe(g_field->one()); }
template <class F> Fe fe_add(const Fe& a, const Fe& b) { return fe(g_field->add(el(a), el(b))); }
template <class F> Fe fe_sub(const Fe& a, const Fe& b) { return fe(g_field->sub(el(a), el(b))); }
template <class F> Fe fe_mul(const Fe& a, const Fe& b) { return fe(g_field->mul(el(a), el(b))); }
template <class F> Fe fe_fold_fixed(const Fe& l, const Fe& h, const FixedMul&) {
    return fe(g_field->sub(el(l), g_field->mul(g_challenge, g_field->sub(el(l), el(h)))));
}
template <class F> void fe_fold_fixed_f64_x2(Fe& lo, Fe& hi, const Fe& x0, const Fe& x1, const Fe& x2, const Fe& x3, const FixedMulF64&) {
    lo = fe_fold_fixed<F>(x0, x2, FixedMul{});
    hi = fe_fold_fixed<F>(x1, x3, FixedMul{});
}
inline Fe ld_fe(const Fe* p) { return *p; }
inline Fe ld_fe_stream(const Fe* p) { return *p; }
inline void st_fe(Fe* p, const Fe& v) { *p = v; }
#include "../host_accw.hpp"
struct ReduceArgs { int skip1; };
// hooks of sop_kernel.cuh's dynamic work distribution: never reached in the sequential replay (DYN = false)
inline uint32_t sop_fetch_chunk(const ReduceArgs&) { std::abort(); }
inline uint32_t sop_bcast_lane0(uint32_t v) { return v; }
template <class F, int NP, bool TOOM = false>  // the mock replays the plain point set (TOOM = false)
void reduce_publish(const Fe* acc, const ReduceArgs&) {
    for (int t = 0; t < NP; t++) g_sums[(size_t)t] = g_field->add(g_sums[(size_t)t], el(acc[t]));
}
}  // namespace
}  // namespace zk

#include "ntt_sharded_kernels.cuh"
#include "sop_kernel.cuh"
#include "sop_group.hpp"

namespace zk {
namespace {

using host::El;
using host::Field;

template <class K>
void replay(unsigned grid, unsigned threads, K kernel) {
    gridDim = dim3(grid, 1, 1);
    blockDim = dim3(threads, 1, 1);
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < threads; t++) {
            blockIdx = uint3{b, 0, 0};
            threadIdx = uint3{t, 0, 0};
            kernel();
        }
}

// what the last block of a reducing launch does (reduce.cuh: reduce_publish)
void publish(const ReduceScratch& s, const std::vector<El>& vals, int slot = 0) {
    for (size_t t = 0; t < vals.size(); t++) {
        const Fe v = fe(vals[t]);
        s.result_dev[slot + (int)t] = v;
        s.result_host_devptr[slot + (int)t] = v;
        if (s.lanes)
            for (int i = 0; i < 8; i++) s.lanes[((size_t)slot + t) * 8 + i] = v.v[i];
    }
    if (s.seq != 0) *s.flag_host_devptr = s.seq;
}

std::vector<El> round_sums(const Field& F, const TablePtrs& tabs, int m, int degree, uint64_t half) {
    std::vector<El> S((size_t)degree + 1, F.zero());
    for (int t = 0; t <= degree; t++) {
        const El ft = F.from_u64((uint64_t)t);
        for (uint64_t j = 0; j < half; j++) {
            El pr = F.one();
            for (int k = 0; k < m; k++) {
                const El lo = el(tabs.t[k][j]), hi = el(tabs.t[k][j + half]);
                pr = F.mul(pr, F.sub(lo, F.mul(ft, F.sub(lo, hi))));
            }
            S[(size_t)t] = F.add(S[(size_t)t], pr);
        }
    }
    return S;
}

void fold_all(const Field& F, const TablePtrs& tabs, int m, uint64_t half, const El& r) {
    for (int k = 0; k < m; k++)
        for (uint64_t j = 0; j < half; j++) {
            const El l = el(tabs.t[k][j]), h = el(tabs.t[k][j + half]);
            tabs.t[k][j] = fe(F.sub(l, F.mul(r, F.sub(l, h))));
        }
}

}  // namespace

struct NttPlan {
    int field;
    unsigned log_n;
    bool inverse;
    Fe* scratch;
};

bool has_fused_path(int m, int degree) { return m >= 1 && m <= kMaxFactors && degree >= 1 && degree <= 4; }
bool sop_degree_supported(int degree) { return degree >= 1 && degree <= 4; }

cudaError_t launch_round_poly(int field, const TablePtrs& tabs, int m, int degree, uint64_t half, const ReduceScratch& s,
                              cudaStream_t, int* launches) {
    const Field F(field);
    ++*launches;
    publish(s, round_sums(F, tabs, m, degree, half));
    return cudaSuccess;
}
cudaError_t launch_round_poly_range(int field, const TablePtrs& tabs, int m, int degree, uint64_t count, uint64_t hoff,
                                    const ReduceScratch& s, cudaStream_t, int* launches) {
    const Field F(field);
    ++*launches;
    std::vector<El> S((size_t)degree + 1, F.zero());
    for (int t = 0; t <= degree; t++) {
        const El ft = F.from_u64((uint64_t)t);
        for (uint64_t j = 0; j < count; j++) {
            El pr = F.one();
            for (int k = 0; k < m; k++) {
                const El lo = el(tabs.t[k][j]), hi = el(tabs.t[k][j + hoff]);
                pr = F.mul(pr, F.sub(lo, F.mul(ft, F.sub(lo, hi))));
            }
            S[(size_t)t] = F.add(S[(size_t)t], pr);
        }
    }
    publish(s, S);
    return cudaSuccess;
}
cudaError_t launch_fold(int field, const TablePtrs& tabs, int m, uint64_t half, const Fe& r, cudaStream_t, int* launches) {
    const Field F(field);
    ++*launches;
    fold_all(F, tabs, m, half, el(r));
    return cudaSuccess;
}
cudaError_t launch_fold_round_poly(int field, const TablePtrs& tabs, int m, int degree, uint64_t n_prev, const Fe& r,
                                   const ReduceScratch& s, cudaStream_t, int* launches, const Fe* claim) {
    const Field F(field);
    ++*launches;
    fold_all(F, tabs, m, n_prev / 2, el(r));
    std::vector<El> S = round_sums(F, tabs, m, degree, n_prev / 4);
    if (claim && degree >= 1) S[1] = F.sub(el(*claim), S[0]);  // the derived S(1): wrong if the host's claim is wrong
    publish(s, S);
    return cudaSuccess;
}
cudaError_t launch_product_sum(int field, const TablePtrs& tabs, int m, uint64_t n, const ReduceScratch& s, cudaStream_t,
                               int* launches) {
    const Field F(field);
    ++*launches;
    El acc = F.zero();
    for (uint64_t j = 0; j < n; j++) {
        El pr = el(tabs.t[0][j]);
        for (int k = 1; k < m; k++) pr = F.mul(pr, el(tabs.t[k][j]));
        acc = F.add(acc, pr);
    }
    publish(s, {acc});
    return cudaSuccess;
}

// ---- sum of products: the real kernel source, replayed
template <class FT, int D>
cudaError_t sop_replay(int field, const TablePtrs& tabs, const SopSpec& spec_in, uint64_t q, bool fold, const Fe& r,
                       const ReduceScratch& s, const Fe* claim) {
    const Field F(field);
    g_field = &F;
    g_challenge = el(r);
    g_sums.assign((size_t)D + 1, F.zero());
    const unsigned grid = 3;
    const int skip1 = (fold && claim) ? 1 : 0;
    const char* wide_env = std::getenv("ZK_B200_SOP_WIDE");  // the same knobs the product launcher reads (kernels_sop.cu)
    const char* group_env = std::getenv("ZK_B200_SOP_GROUP");
    const SopSpec spec = (group_env && group_env[0] == '0') ? spec_in : sop_group(spec_in);  // common factors, like the launcher
    if (wide_env && wide_env[0] == '1') {
        // the deferred-reduction variant: block by block, shared memory cleared first (the kernel's accw_zero +
        // __syncthreads(), which a thread-by-thread replay cannot interleave)
        for (unsigned b = 0; b < grid; b++) {
            std::memset(sop_smem, 0, sizeof(sop_smem));
            gridDim = dim3(grid, 1, 1);
            blockDim = dim3(kThreads, 1, 1);
            for (unsigned t = 0; t < (unsigned)kThreads; t++) {
                blockIdx = uint3{b, 0, 0};
                threadIdx = uint3{t, 0, 0};
                if (fold) sop_round_kernel<FT, D, true, false, true>(tabs, spec, q, FixedMul{}, FixedMulF64Sel{}, ReduceArgs{skip1});
                else sop_round_kernel<FT, D, false, false, true>(tabs, spec, q, FixedMul{}, FixedMulF64Sel{}, ReduceArgs{0});
            }
        }
    } else if (fold) replay(grid, kThreads, [&] { sop_round_kernel<FT, D, true, false, false>(tabs, spec, q, FixedMul{}, FixedMulF64Sel{}, ReduceArgs{skip1}); });
    else replay(grid, kThreads, [&] { sop_round_kernel<FT, D, false, false, false>(tabs, spec, q, FixedMul{}, FixedMulF64Sel{}, ReduceArgs{0}); });
    if (skip1) g_sums[1] = F.sub(el(*claim), g_sums[0]);
    publish(s, g_sums);
    g_field = nullptr;
    return cudaSuccess;
}
template <class FT>
cudaError_t sop_deg(int field, const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t q, bool fold, const Fe& r,
                    const ReduceScratch& s, const Fe* claim) {
    switch (degree) {
        case 1: return sop_replay<FT, 1>(field, tabs, spec, q, fold, r, s, claim);
        case 2: return sop_replay<FT, 2>(field, tabs, spec, q, fold, r, s, claim);
        case 3: return sop_replay<FT, 3>(field, tabs, spec, q, fold, r, s, claim);
        case 4: return sop_replay<FT, 4>(field, tabs, spec, q, fold, r, s, claim);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t launch_sop_round_poly(int field, const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t half,
                                  const ReduceScratch& s, cudaStream_t, int* launches) {
    ++*launches;
    return field == 0 ? sop_deg<Fr381>(field, tabs, spec, degree, half, false, Fe{}, s, nullptr)
                      : sop_deg<Fr377>(field, tabs, spec, degree, half, false, Fe{}, s, nullptr);
}
cudaError_t launch_sop_fold_round_poly(int field, const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t n_prev,
                                       const Fe& r, const ReduceScratch& s, cudaStream_t, int* launches, const Fe* claim) {
    ++*launches;
    return field == 0 ? sop_deg<Fr381>(field, tabs, spec, degree, n_prev / 4, true, r, s, claim)
                      : sop_deg<Fr377>(field, tabs, spec, degree, n_prev / 4, true, r, s, claim);
}

// ---- MLE utilities: naive models of the contracts in kernels.h
cudaError_t launch_fold_var(int field, const Fe* in, Fe* out, unsigned nv, unsigned initial_var, const Fe& a, cudaStream_t,
                            int* launches) {
    const Field F(field);
    ++*launches;
    const unsigned pos = nv - 1 - initial_var;
    const uint64_t pairs = (uint64_t)1 << (nv - 1), low = ((uint64_t)1 << pos) - 1;
    std::vector<Fe> tmp(pairs);
    for (uint64_t k = 0; k < pairs; k++) {
        const uint64_t l = ((k >> pos) << (pos + 1)) | (k & low), r = l | ((uint64_t)1 << pos);
        tmp[k] = fe(F.sub(el(in[l]), F.mul(el(a), F.sub(el(in[l]), el(in[r])))));
    }
    for (uint64_t k = 0; k < pairs; k++) out[k] = tmp[k];
    return cudaSuccess;
}
cudaError_t launch_prod_reduce(int field, const TablePtrs& tabs, int m, uint64_t n, Fe* out, cudaStream_t, int* launches) {
    const Field F(field);
    ++*launches;
    for (uint64_t j = 0; j < n; j++) {
        El pr = el(tabs.t[0][j]);
        for (int k = 1; k < m; k++) pr = F.mul(pr, el(tabs.t[k][j]));
        out[j] = fe(pr);
    }
    return cudaSuccess;
}
cudaError_t launch_to_bytes(int field, const Fe* in, uint64_t n, uint8_t* out, cudaStream_t, int* launches) {
    const Field F(field);
    ++*launches;
    for (uint64_t j = 0; j < n; j++) F.to_be32(el(in[j]), out + 32 * j);
    return cudaSuccess;
}
cudaError_t launch_convert(int, Fe*, uint64_t, bool, cudaStream_t, int*) { return cudaErrorNotSupported; }
static inline uint64_t splitmix64_mix(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
    return z;
}
cudaError_t launch_generate(int field, Fe* out, uint64_t count, uint64_t seed, uint64_t table_id, uint64_t first,
                            uint64_t stride, cudaStream_t, int* launches) {
    const Field F(field);
    ++*launches;
    for (uint64_t j = 0; j < count; j++) {
        const uint64_t i = first + j * stride;
        uint64_t c[4];
        for (int l = 0; l < 4; l++) c[l] = splitmix64_mix(seed + 0x9E3779B97F4A7C15ULL * ((((table_id << 40) + i) * 4 + (uint64_t)l) + 1));
        c[3] &= 0x3FFFFFFFFFFFFFFFULL;
        while (F.geq_p(c)) F.sub_p(c);
        out[j] = fe(F.from_canonical(c));
    }
    return cudaSuccess;
}
cudaError_t launch_interleave(const Fe* in, Fe* out, uint64_t local_len, unsigned world, cudaStream_t, int* launches, unsigned batch) {
    ++*launches;
    const uint64_t per = local_len * world;
    for (uint64_t b = 0; b < batch; b++)
        for (uint64_t q = 0; q < world; q++)
            for (uint64_t j = 0; j < local_len; j++) out[b * per + j * world + q] = in[b * per + q * local_len + j];
    return cudaSuccess;
}
cudaError_t launch_deinterleave(const Fe* in, Fe* out, uint64_t local_len, unsigned world, cudaStream_t, int* launches) {
    ++*launches;
    for (uint64_t q = 0; q < world; q++)
        for (uint64_t j = 0; j < local_len; j++) out[q * local_len + j] = in[j * world + q];
    return cudaSuccess;
}
// after the exact u64 all-reduce of one 32-bit limb per lane: carry-propagate, reduce mod p, publish result + flag
cudaError_t launch_narrow(int field, const uint64_t* lanes, Fe* out_dev, Fe* out_host_devptr, int count, unsigned* flag_host_devptr,
                          unsigned seq, cudaStream_t, int* launches) {
    const Field F(field);
    ++*launches;
    for (int t = 0; t < count; t++) {
        // value = sum_i lanes[i] * 2^(32 i), up to ~2^(256+3): reduce by repeated subtraction of p on a 5-word number
        uint64_t w[5] = {0, 0, 0, 0, 0};
        unsigned __int128 carry = 0;
        uint32_t limbs[10] = {0};
        for (int i = 0; i < 8; i++) {
            carry += lanes[(size_t)t * 8 + i];
            limbs[i] = (uint32_t)carry;
            carry >>= 32;
        }
        limbs[8] = (uint32_t)carry;
        limbs[9] = (uint32_t)(carry >> 32);
        for (int i = 0; i < 5; i++) w[i] = (uint64_t)limbs[2 * i] | ((uint64_t)limbs[2 * i + 1] << 32);
        const host::FieldParams& P = host::params(field);
        auto geq = [&]() {
            if (w[4]) return true;
            for (int i = 3; i >= 0; i--)
                if (w[i] != P.p[i]) return w[i] > P.p[i];
            return true;
        };
        while (geq()) {
            unsigned __int128 borrow = 0;
            for (int i = 0; i < 5; i++) {
                const unsigned __int128 d = (unsigned __int128)w[i] - (i < 4 ? P.p[i] : 0) - (uint64_t)borrow;
                w[i] = (uint64_t)d;
                borrow = (d >> 64) & 1;
            }
        }
        Fe v;
        std::memcpy(v.v, w, 32);
        out_dev[t] = v;
        out_host_devptr[t] = v;
    }
    if (seq != 0) *flag_host_devptr = seq;
    (void)F;
    return cudaSuccess;
}

// ---- single-GPU NTT: a plain iterative transform with the contract of kernels_ntt.cu, INCLUDING where the result
// lands (sizes >= 2^6 end up in the plan's scratch buffer, the caller swaps or copies) so that the callers' buffer
// bookkeeping is exercised
cudaError_t ntt_plan_create(int field, unsigned log_n, bool inverse, cudaStream_t, NttPlan** out, int*) {
    *out = new NttPlan{field, log_n, inverse, nullptr};
    return cudaSuccess;
}
void ntt_plan_destroy(NttPlan* p) {
    if (!p) return;
    std::free(p->scratch);
    delete p;
}
bool ntt_plan_is(const NttPlan* p, int field, unsigned log_n, bool inverse) {
    return p->field == field && p->log_n == log_n && p->inverse == inverse;
}
void ntt_plan_adopt_scratch(NttPlan* plan, Fe* buf) { plan->scratch = buf; }
cudaError_t ntt_execute(NttPlan* p, Fe* data, Fe** result, cudaStream_t, int* launches) {
    const Field F(p->field);
    ++*launches;
    const uint64_t n = (uint64_t)1 << p->log_n;
    std::vector<El> a(n);
    for (uint64_t i = 0; i < n; i++) {  // bit-reversed load
        uint64_t r = 0;
        for (unsigned b = 0; b < p->log_n; b++) r |= ((i >> b) & 1) << (p->log_n - 1 - b);
        a[r] = el(data[i]);
    }
    El w_n = F.root_of_unity(p->log_n);
    if (p->inverse) w_n = F.inverse(w_n);
    for (unsigned s = 1; s <= p->log_n; s++) {
        const uint64_t len = (uint64_t)1 << s, half = len / 2;
        const uint64_t e[1] = {n / len};
        const El w_len = F.pow(w_n, e, 1);
        for (uint64_t b = 0; b < n; b += len) {
            El w = F.one();
            for (uint64_t j = 0; j < half; j++) {
                const El u = a[b + j], v = F.mul(a[b + j + half], w);
                a[b + j] = F.add(u, v);
                a[b + j + half] = F.sub(u, v);
                w = F.mul(w, w_len);
            }
        }
    }
    if (p->inverse) {
        const El n_inv = F.inverse(F.from_u64(n));
        for (auto& x : a) x = F.mul(x, n_inv);
    }
    Fe* dst = data;
    if (p->log_n >= 6) {
        if (!p->scratch) p->scratch = (Fe*)std::aligned_alloc(256, (n * sizeof(Fe) + 255) / 256 * 256);
        dst = p->scratch;
        for (uint64_t i = 0; i < n; i++) data[i] = fe(F.zero());  // `data` is clobbered
    }
    for (uint64_t i = 0; i < n; i++) dst[i] = fe(a[i]);
    *result = dst;
    return cudaSuccess;
}

// ---- multi-GPU NTT pieces: the real kernel sources, replayed
template <class FT>
void sharded_pow_table(Fe* out, uint64_t count, const Fe& base, unsigned shift) {
    replay(2, kShThreads, [&] { pow_table_kernel<FT>(out, count, base, shift); });
}
cudaError_t launch_pow_table(int field, Fe* out, uint64_t count, const Fe& base, unsigned shift, cudaStream_t, int* launches) {
    const Field F(field);
    g_field = &F;
    ++*launches;
    if (field == 0) sharded_pow_table<Fr381>(out, count, base, shift); else sharded_pow_table<Fr377>(out, count, base, shift);
    g_field = nullptr;
    return cudaSuccess;
}
cudaError_t launch_twiddle_mul(int field, Fe* x, uint64_t m, const Fe* t_lo, const Fe* t_hi, unsigned lo_bits, cudaStream_t,
                               int* launches) {
    const Field F(field);
    g_field = &F;
    ++*launches;
    if (field == 0) replay(3, kShThreads, [&] { twiddle_mul_kernel<Fr381>(x, m, t_lo, t_hi, lo_bits); });
    else replay(3, kShThreads, [&] { twiddle_mul_kernel<Fr377>(x, m, t_lo, t_hi, lo_bits); });
    g_field = nullptr;
    return cudaSuccess;
}
template <class FT>
cudaError_t gdft_replay(int ranks, const Fe* in, Fe* out, uint64_t chunk, const GdftParams& prm) {
    switch (ranks) {
        case 2: replay(2, kShThreads, [&] { gdft_kernel<FT, 2>(in, out, chunk, prm); }); break;
        case 4: replay(2, kShThreads, [&] { gdft_kernel<FT, 4>(in, out, chunk, prm); }); break;
        case 8: replay(2, kShThreads, [&] { gdft_kernel<FT, 8>(in, out, chunk, prm); }); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}
cudaError_t launch_gdft(int field, int ranks, const Fe* in, Fe* out, uint64_t chunk, const Fe* w_half, const Fe* scale,
                        cudaStream_t, int* launches) {
    const Field F(field);
    g_field = &F;
    ++*launches;
    GdftParams prm{};
    for (int i = 0; i < ranks / 2; i++) prm.w[i] = w_half[i];
    prm.do_scale = scale ? 1 : 0;
    if (scale) prm.scale = *scale;
    cudaError_t e = field == 0 ? gdft_replay<Fr381>(ranks, in, out, chunk, prm) : gdft_replay<Fr377>(ranks, in, out, chunk, prm);
    g_field = nullptr;
    return e;
}

cudaError_t run_microbench(int, MicrobenchResult*, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace zk
