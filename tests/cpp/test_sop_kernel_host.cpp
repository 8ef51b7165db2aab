// Host replay of the sum-of-products round kernel (zk_b200/csrc/sop_kernel.cuh), CPU suite.
//
// The kernel source is compiled here as plain C++: the CUDA decorations expand to nothing outside nvcc, the device
// field arithmetic (inline PTX in field.cuh, compiled out without __CUDACC__) is replaced by host_field.hpp's
// word-serial Montgomery arithmetic, shared memory is an ordinary array, and the grid is replayed block by block,
// thread by thread.  What this pins without a GPU: the pair / quadruple index conventions of the fused in-place
// fold (evaluation_form.rs:40-80 with initial_var = 0), the term bookkeeping, the evaluation points t = 0..D, the
// grid-stride loop with ragged sizes, the deferred-reduction (WIDE) variant's shared-memory carve-up and term handling.  What it cannot pin (the PTX multiplier, the block/grid reduction) is shared
// with the product kernels and covered by the GPU suite.
// The expected values come from an independent, deliberately naive model written below: fold out of place, then
// evaluate every term at every t from scratch with multiplications by t.
#include <cuda_runtime.h>  // vector types only (plain g++)

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define __launch_bounds__(...)
static uint3 threadIdx, blockIdx;
static dim3 gridDim, blockDim;

#include "kernels.h"
#include "field_f64.cuh"
#include "host_field.hpp"

namespace zk {
namespace {

constexpr int kThreads = 128;
const host::Field* g_field = nullptr;
host::El g_challenge;                       // the launch-wide fold multiplier (the device reads it from FixedMul)
std::vector<host::El> g_sums;               // what reduce_publish would publish
// shared memory of one block, sized by the harness exactly as the launcher does (kernels_sop.cu: sop_smem_total) and
// followed by guard words: a carve-up that runs past the launch's allocation is caught
constexpr size_t kSmemMaxUint4 = 2 * 2 * (kMaxFactors + kMaxVirtual) * kThreads + 5 * (4 * kThreads + kThreads / 4) + 64;
alignas(32) uint4 sop_smem[kSmemMaxUint4];

inline host::El el(const Fe& a) { host::El e; std::memcpy(e.v, a.v, 32); return e; }
inline Fe fe(const host::El& e) { Fe a; std::memcpy(a.v, e.v, 32); return a; }
template <class F> Fe fe_zero() { return fe(g_field->zero()); }
template <class F> Fe fe_add(const Fe& a, const Fe& b) { return fe(g_field->add(el(a), el(b))); }
template <class F> Fe fe_sub(const Fe& a, const Fe& b) { return fe(g_field->sub(el(a), el(b))); }
template <class F> Fe fe_mul(const Fe& a, const Fe& b) { return fe(g_field->mul(el(a), el(b))); }
template <class F> Fe fe_fold_fixed(const Fe& l, const Fe& h, const FixedMul&) {  // l - r (l - h)
    return fe(g_field->sub(el(l), g_field->mul(g_challenge, g_field->sub(el(l), el(h)))));
}
inline Fe ld_fe_stream(const Fe* p) { return *p; }
inline void st_fe(Fe* p, const Fe& v) { *p = v; }
template <class F> void fe_fold_fixed_f64_x2(Fe& lo, Fe& hi, const Fe& x0, const Fe& x1, const Fe& x2, const Fe& x3,
                                             const FixedMulF64&) {  // the FP64-pipe variant computes the same two folds
    lo = fe_fold_fixed<F>(x0, x2, FixedMul{});
    hi = fe_fold_fixed<F>(x1, x3, FixedMul{});
}
#include "host_accw.hpp"
struct ReduceArgs { int skip1; };
// hooks of the dynamic work distribution: never reached in the sequential replay (DYN = false)
inline uint32_t sop_fetch_chunk(const ReduceArgs&) { std::abort(); }
inline uint32_t sop_bcast_lane0(uint32_t v) { return v; }
template <class F, int NP, bool TOOM = false>
void reduce_publish(const Fe* acc, const ReduceArgs&) {  // the harness applies S(1) = claim - S(0) (and the Toom map) after the last block
    for (int t = 0; t < NP; t++) g_sums[(size_t)t] = g_field->add(g_sums[(size_t)t], el(acc[t]));
}

}  // namespace
}  // namespace zk

#include "sop_kernel.cuh"
#include "sop_group.hpp"

using zk::Fe;
using zk::host::El;
using zk::host::Field;

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static El rnd_el(const Field& F) {  // a product of two random words times a third: spread over the whole field
    El a = F.from_u64(rnd()), b = F.from_u64(rnd()), c = F.from_u64(rnd());
    return F.add(F.mul(F.mul(a, b), F.mul(c, b)), a);
}

static long g_guard_hits = 0;
template <class FT, int D, bool FOLD, bool F64, bool WIDE, bool TOOM = false>
static void replay(const zk::TablePtrs& tabs, const zk::SopSpec& spec, uint64_t q, unsigned grid, int skip1) {
    gridDim = dim3(grid, 1, 1);
    blockDim = dim3(zk::kThreads, 1, 1);
    // what the launcher allocates for this launch, in uint4 units
    const size_t used = ((size_t)2 * (spec.n_tables + spec.n_virt) * zk::kThreads * sizeof(zk::Fe) + (WIDE ? zk::accw_bytes(D + 1) : 0)) / sizeof(uint4);
    for (unsigned b = 0; b < grid; b++) {
        // a block starts with whatever the previous one left behind, except that the kernel's own accw_zero +
        // __syncthreads() happen before any accumulation: in this sequential replay that is a clear before the block
        for (size_t i = 0; i < zk::kSmemMaxUint4; i++) zk::sop_smem[i] = (i < used) ? uint4{0, 0, 0, 0} : uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu};
        for (unsigned t = 0; t < (unsigned)zk::kThreads; t++) {
            blockIdx = uint3{b, 0, 0};
            threadIdx = uint3{t, 0, 0};
            zk::sop_round_kernel<FT, D, FOLD, F64, WIDE, false, TOOM>(tabs, spec, q, zk::FixedMul{}, zk::FixedMulF64Sel{}, zk::ReduceArgs{skip1});
        }
        for (size_t i = used; i < zk::kSmemMaxUint4; i++) g_guard_hits += (zk::sop_smem[i].x != 0xDEADBEEFu || zk::sop_smem[i].w != 0xDEADBEEFu);
    }
}

template <class FT, int D>
static long run_case(int field, unsigned log_len, bool fold, const zk::SopSpec& spec, unsigned grid, int mode) {
    const Field F(field);
    zk::g_field = &F;
    const size_t len = (size_t)1 << log_len;  // table length entering the launch
    std::vector<std::vector<El>> T((size_t)spec.n_tables, std::vector<El>(len));
    for (auto& t : T)
        for (auto& e : t) e = rnd_el(F);
    // edge values in the first table: 0, 1, p-1
    T[0][0] = F.zero();
    T[0][len - 1] = F.one();
    if (len > 2) T[0][1] = F.sub(F.zero(), F.one());
    const El r = rnd_el(F);
    zk::g_challenge = r;

    // naive model
    std::vector<std::vector<El>> M = T;
    size_t cur = len;
    if (fold) {
        cur = len / 2;
        for (auto& t : M) {
            std::vector<El> nx(cur);
            for (size_t j = 0; j < cur; j++) nx[j] = F.sub(t[j], F.mul(r, F.sub(t[j], t[j + cur])));  // pair (j, j + N/2)
            t = nx;
        }
    }
    const size_t half = cur / 2;
    std::vector<El> want((size_t)D + 1, F.zero());
    for (int t = 0; t <= D; t++) {
        const El ft = F.from_u64((uint64_t)t);
        for (size_t j = 0; j < half; j++)
            for (int term = 0; term < spec.n_terms; term++) {
                El pr = F.one();
                for (int i = 0; i < (int)spec.len[term]; i++) {
                    const std::vector<El>& A = M[spec.fac[term][i]];
                    pr = F.mul(pr, F.sub(A[j], F.mul(ft, F.sub(A[j], A[j + half]))));  // left - t (left - right)
                }
                want[(size_t)t] = F.add(want[(size_t)t], pr);
            }
    }

    // the kernel source, replayed
    std::vector<std::vector<Fe>> dev((size_t)spec.n_tables, std::vector<Fe>(len));
    zk::TablePtrs tabs{};
    for (int k = 0; k < spec.n_tables; k++) {
        for (size_t j = 0; j < len; j++) dev[(size_t)k][j] = zk::fe(T[(size_t)k][j]);
        tabs.t[k] = dev[(size_t)k].data();
    }
    zk::g_sums.assign((size_t)D + 1, F.zero());
    // the kernel runs the launcher's GROUPED spec (common factors through virtual tables, sop_group.hpp); the naive model
    // above evaluated the original terms
    const zk::SopSpec kspec = zk::sop_group(spec);
    // mode 0: every evaluation summed, integer folds; 1: S(1) derived from the claim; 2: S(1) derived, FP64 folds;
    // 3: deferred reduction (WIDE), every evaluation summed; 4: WIDE with S(1) derived;
    // 5 / 6 (D == 3, no term longer than 3): the Toom point set (0, 1, -1, infinity), all summed / S(1) derived
    bool cubic_at_most = (D == 3);
    for (int t = 0; t < spec.n_terms; t++) cubic_at_most = cubic_at_most && spec.len[t] <= 3;
    const bool toom = mode >= 5;
    if (toom && !cubic_at_most) return 0;
    if (toom) {
        constexpr bool T = (D == 3);
        if (fold) replay<FT, D, true, false, false, T>(tabs, kspec, len / 4, grid, mode == 6);
        else replay<FT, D, false, false, false, T>(tabs, kspec, len / 2, grid, 0);
    } else
    if (fold && mode == 2) replay<FT, D, true, true, false>(tabs, kspec, len / 4, grid, 1);
    else if (fold && mode >= 3) replay<FT, D, true, false, true>(tabs, kspec, len / 4, grid, mode == 4);
    else if (fold) replay<FT, D, true, false, false>(tabs, kspec, len / 4, grid, mode == 1);
    else if (mode >= 3) replay<FT, D, false, false, true>(tabs, kspec, len / 2, grid, 1);
    else replay<FT, D, false, false, false>(tabs, kspec, len / 2, grid, 1 /* ignored without a fold */);
    if (fold && (mode == 1 || mode == 2 || mode == 4 || mode == 6)) {  // what the last block does with ra.claim = S(0) + S(1)
        long untouched = (zk::g_sums[1] != F.zero());
        zk::g_sums[1] = F.sub(F.add(want[0], want[1]), zk::g_sums[0]);
        if (untouched) return 1000000;  // the t = 1 products were not skipped
    }

    if (toom) {  // reduce.cuh: toom_to_evals — {S(0), S(1), S(-1), c3} -> S(0..3)
        const El half = F.inverse(F.from_u64(2));
        const El s0 = zk::g_sums[0], s1 = zk::g_sums[1], sm = zk::g_sums[2], c3 = zk::g_sums[3];
        const El c2 = F.sub(F.mul(F.add(s1, sm), half), s0), c1 = F.sub(F.mul(F.sub(s1, sm), half), c3);
        zk::g_sums[2] = F.add(F.add(s0, F.mul(F.from_u64(2), c1)), F.add(F.mul(F.from_u64(4), c2), F.mul(F.from_u64(8), c3)));
        zk::g_sums[3] = F.add(F.add(s0, F.mul(F.from_u64(3), c1)), F.add(F.mul(F.from_u64(9), c2), F.mul(F.from_u64(27), c3)));
    }
    long bad = 0;
    for (int t = 0; t <= D; t++) bad += (zk::g_sums[(size_t)t] != want[(size_t)t]);
    if (fold)  // the folded tables are the first half of the buffers, in place
        for (int k = 0; k < spec.n_tables; k++)
            for (size_t j = 0; j < cur; j++) bad += (zk::el(dev[(size_t)k][j]) != M[(size_t)k][j]);
    return bad;
}

static zk::SopSpec make_spec(int n_tables, std::vector<std::vector<int>> terms) {
    zk::SopSpec s{};
    s.n_tables = n_tables;
    s.n_terms = (int)terms.size();
    for (size_t t = 0; t < terms.size(); t++) {
        s.len[t] = (uint8_t)terms[t].size();
        for (size_t i = 0; i < terms[t].size(); i++) s.fac[t][i] = (uint8_t)terms[t][i];
    }
    return s;
}

int main() {
    long bad = 0, cases = 0;
    const zk::SopSpec gkr = make_spec(4, {{0, 2}, {0, 3}, {1, 2, 3}});
    const zk::SopSpec sq = make_spec(2, {{0, 0}, {1}});
    const zk::SopSpec wide = make_spec(8, {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 7}, {1, 6}, {2, 5}, {3, 4}});
    const zk::SopSpec one = make_spec(3, {{0, 1, 2}});
    // shapes the launcher factors: x.a + x.b -> x.(a + b)
    const zk::SopSpec fan = make_spec(4, {{0, 1}, {0, 2}, {0, 3}});
    const zk::SopSpec deep = make_spec(4, {{0, 1, 2}, {1, 0, 3}, {1, 2}, {3, 3}});
    {
        const zk::SopSpec g = zk::sop_group(gkr), f = zk::sop_group(fan), d = zk::sop_group(deep), w = zk::sop_group(wide);
        if (g.n_terms != 2 || g.n_virt != 1 || f.n_terms != 2 || f.n_virt != 1 || d.n_terms != 3 || d.n_virt != 1 || w.n_virt != 4 || w.n_terms != 4) {
            std::printf("sop_group: unexpected factoring (%d,%d) (%d,%d) (%d,%d) (%d,%d)\n", g.n_terms, g.n_virt, f.n_terms, f.n_virt, d.n_terms, d.n_virt, w.n_terms, w.n_virt);
            return 1;
        }
    }
    for (int field = 0; field < 2; field++) {
        for (unsigned log_len = 1; log_len <= 11; log_len++) {
            for (int fold = 0; fold < 2; fold++) {
                if (fold && log_len < 2) continue;  // the fused launch needs 4 entries
                for (unsigned grid : {1u, 3u})
                    for (int mode = 0; mode < 7; mode++) {
                        if (!fold && mode != 0 && mode != 3 && mode != 5) continue;
                        if (field == 0) {
                            bad += run_case<zk::Fr381, 3>(field, log_len, fold, gkr, grid, mode);
                            bad += run_case<zk::Fr381, 2>(field, log_len, fold, sq, grid, mode);
                            bad += run_case<zk::Fr381, 2>(field, log_len, fold, wide, grid, mode);
                            bad += run_case<zk::Fr381, 4>(field, log_len, fold, one, grid, mode);
                            bad += run_case<zk::Fr381, 1>(field, log_len, fold, sq, grid, mode);
                            bad += run_case<zk::Fr381, 2>(field, log_len, fold, fan, grid, mode);
                            bad += run_case<zk::Fr381, 3>(field, log_len, fold, deep, grid, mode);
                        } else {
                            bad += run_case<zk::Fr377, 3>(field, log_len, fold, gkr, grid, mode);
                            bad += run_case<zk::Fr377, 2>(field, log_len, fold, sq, grid, mode);
                            bad += run_case<zk::Fr377, 2>(field, log_len, fold, wide, grid, mode);
                            bad += run_case<zk::Fr377, 4>(field, log_len, fold, one, grid, mode);
                            bad += run_case<zk::Fr377, 1>(field, log_len, fold, sq, grid, mode);
                            bad += run_case<zk::Fr377, 2>(field, log_len, fold, fan, grid, mode);
                            bad += run_case<zk::Fr377, 3>(field, log_len, fold, deep, grid, mode);
                        }
                        cases += 7;
                    }
            }
        }
    }
    std::printf("%ld cases, %ld mismatches, %ld guard hits\n", cases, bad, g_guard_hits);
    return (bad || g_guard_hits) ? 1 : 0;
}
