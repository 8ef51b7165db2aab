// Host replay of the multi-GPU NTT's own kernels (zk_b200/csrc/ntt_sharded_kernels.cuh), CPU suite: the power table,
// the inter-rank twiddle multiplication and the G-point DFT across ranks, compiled as plain C++ with host_field.hpp
// standing in for the device arithmetic, replayed thread by thread and compared with direct formulas
// (x * w^k by repeated multiplication; the naive O(G^2) DFT).  The whole factorisation is checked separately over
// Python integers (tests/test_ntt_sharded_model.py); the local M-point transform is the single-GPU NTT of the GPU suite.
#include <cuda_runtime.h>  // vector types only (plain g++)

#include <cstdio>
#include <cstring>
#include <vector>

#define __launch_bounds__(...)
static uint3 threadIdx, blockIdx;
static dim3 gridDim, blockDim;

#include "kernels.h"
#include "host_field.hpp"

namespace zk {
namespace {
const host::Field* g_field = nullptr;
inline host::El el(const Fe& a) { host::El e; std::memcpy(e.v, a.v, 32); return e; }
inline Fe fe(const host::El& e) { Fe a; std::memcpy(a.v, e.v, 32); return a; }
template <class F> Fe fe_one() { return fe(g_field->one()); }
template <class F> Fe fe_add(const Fe& a, const Fe& b) { return fe(g_field->add(el(a), el(b))); }
template <class F> Fe fe_sub(const Fe& a, const Fe& b) { return fe(g_field->sub(el(a), el(b))); }
template <class F> Fe fe_mul(const Fe& a, const Fe& b) { return fe(g_field->mul(el(a), el(b))); }
inline Fe ld_fe(const Fe* p) { return *p; }
inline void st_fe(Fe* p, const Fe& v) { *p = v; }
}  // namespace
}  // namespace zk

#include "ntt_sharded_kernels.cuh"

using zk::Fe;
using zk::host::El;
using zk::host::Field;

static uint64_t rng_state = 12345;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static El rnd_el(const Field& F) {
    El a = F.from_u64(rnd()), b = F.from_u64(rnd()), c = F.from_u64(rnd());
    return F.add(F.mul(F.mul(a, b), F.mul(c, b)), a);
}
template <class K>
static void replay(unsigned grid, K kernel) {
    gridDim = dim3(grid, 1, 1);
    blockDim = dim3(zk::kShThreads, 1, 1);
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < (unsigned)zk::kShThreads; t++) {
            blockIdx = uint3{b, 0, 0};
            threadIdx = uint3{t, 0, 0};
            kernel();
        }
}

template <class FT>
static long check_field(int field) {
    const Field F(field);
    zk::g_field = &F;
    long bad = 0;
    // ---- power tables and the twiddle multiplication: x[k] * w^k, ragged sizes, several splits
    for (unsigned log_m : {0u, 1u, 5u, 9u, 11u}) {
        for (unsigned lo_bits : {0u, 3u, 13u}) {
            if (lo_bits > log_m) continue;
            const uint64_t m = (uint64_t)1 << log_m, n_lo = (uint64_t)1 << lo_bits, n_hi = m >> lo_bits;
            const El w = rnd_el(F);
            std::vector<Fe> t_lo(n_lo), t_hi(n_hi), x(m);
            const Fe wf = zk::fe(w);
            replay(2, [&] { zk::pow_table_kernel<FT>(t_lo.data(), n_lo, wf, 0); });
            replay(1, [&] { zk::pow_table_kernel<FT>(t_hi.data(), n_hi, wf, lo_bits); });
            std::vector<El> x0(m);
            for (uint64_t k = 0; k < m; k++) { x0[k] = rnd_el(F); x[k] = zk::fe(x0[k]); }
            replay(3, [&] { zk::twiddle_mul_kernel<FT>(x.data(), m, t_lo.data(), t_hi.data(), lo_bits); });
            El wk = F.one();
            for (uint64_t k = 0; k < m; k++) {
                bad += (zk::el(x[k]) != F.mul(x0[k], wk));
                wk = F.mul(wk, w);
            }
        }
    }
    return bad;
}

// ---- G-point DFT across ranks, forward and inverse-with-scale, against the O(G^2) definition
template <class FT, int G>
static long check_gdft(int field) {
    const Field F(field);
    zk::g_field = &F;
    long bad = 0;
    El wG = F.root_of_unity(G == 2 ? 1u : (G == 4 ? 2u : 3u));
    for (int inverse = 0; inverse < 2; inverse++) {
        const El w = inverse ? F.inverse(wG) : wG;
        const El scale = F.inverse(F.from_u64((uint64_t)G));
        zk::GdftParams prm{};
        El p = F.one();
        for (int i = 0; i < G / 2; i++) { prm.w[i] = zk::fe(p); p = F.mul(p, w); }
        prm.scale = zk::fe(scale);
        prm.do_scale = inverse;
        for (uint64_t chunk : {(uint64_t)1, (uint64_t)7, (uint64_t)300}) {
            std::vector<Fe> in((size_t)G * chunk), out((size_t)G * chunk);
            std::vector<El> in0((size_t)G * chunk);
            for (size_t i = 0; i < in.size(); i++) { in0[i] = rnd_el(F); in[i] = zk::fe(in0[i]); }
            replay(2, [&] { zk::gdft_kernel<FT, G>(in.data(), out.data(), chunk, prm); });
            for (uint64_t j = 0; j < chunk; j++)
                for (int c = 0; c < G; c++) {
                    El acc = F.zero();
                    for (int q = 0; q < G; q++) {
                        El wp = F.one();
                        for (int e = 0; e < (q * c) % G; e++) wp = F.mul(wp, w);
                        acc = F.add(acc, F.mul(wp, in0[(size_t)q * chunk + j]));
                    }
                    if (inverse) acc = F.mul(acc, scale);
                    bad += (zk::el(out[(size_t)c * chunk + j]) != acc);
                }
        }
    }
    return bad;
}

int main() {
    long bad = 0;
    bad += check_field<zk::Fr381>(0);
    bad += check_field<zk::Fr377>(1);
    bad += check_gdft<zk::Fr381, 2>(0) + check_gdft<zk::Fr381, 4>(0) + check_gdft<zk::Fr381, 8>(0);
    bad += check_gdft<zk::Fr377, 2>(1) + check_gdft<zk::Fr377, 4>(1) + check_gdft<zk::Fr377, 8>(1);
    std::printf("sharded NTT kernels on the host: %ld mismatches\n", bad);
    return bad ? 1 : 0;
}
