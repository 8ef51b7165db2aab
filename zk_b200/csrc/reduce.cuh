// reduce.cuh — what every reducing kernel of the sumcheck path shares (kernels_sumcheck.cu, kernels_sop.cu):
// block size, the warp -> block -> grid reduction of field elements that ends in the last block publishing the
// D+1 evaluations (device buffer, mapped pinned host memory, all-reduce lanes), and the host-side launch helpers.
// Integer modular sums are order independent, so any reduction tree is bit-exact with the reference's sequential
// `.sum()` (sumcheck/src/prover.rs:53-54).
#pragma once
#include <atomic>
#include <cstring>

#include "field_f64.cuh"
#include "host_field.hpp"
#include "kernels.h"

namespace zk {
namespace {

constexpr int kThreads = 128;

constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ Fe ld_fe_cg(const Fe* p) {  // L2-coherent load (other blocks' partials)
    Fe r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                   "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}

template <class F>
__device__ __forceinline__ Fe warp_sum(Fe v, int width = 32) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (off < width) {
            Fe o;
#pragma unroll
            for (int i = 0; i < 8; i++) o.v[i] = __shfl_xor_sync(0xffffffffu, v.v[i], off);
            v = fe_add<F>(v, o);
        }
    }
    return v;
}

struct ReduceArgs {
    Fe* block_partials;
    unsigned* ticket;
    Fe* result_dev;
    Fe* result_host;
    int out_slot;
    unsigned* flag_host;  // mapped pinned word the host spins on (saves a stream synchronisation per round)
    unsigned seq;
    uint64_t* lanes;      // sharded runs: one 32-bit limb per u64 lane, the input of the exact ncclSum all-reduce
    // Rounds >= 1 of a proof: S(0) + S(1) equals the previous round polynomial at its challenge — an identity of the
    // tables, whatever sum the caller claimed — so the kernel skips the products of the t = 1 term and the last block
    // publishes S(1) = claim - S(0): the same field element (prover.rs:49-56 computes it directly).
    int skip1;
    Fe claim;             // this rank's share of S_prev(r_prev): the value on rank 0, zero elsewhere (the map is linear)
    MailboxArgs mbox;     // world > 1: all-reduce through the peer mailboxes inside this launch (kernels.h)
};
constexpr int kWorkCounterOffset = 32;  // the work counter lives 128 bytes after the ticket (own cache line)

// Block-level reduction of NP per-thread accumulators, then grid-level via last-block-done.
// x / 2 in the field (works on any residue representation): (x + (x odd ? p : 0)) >> 1
template <class F>
__device__ __forceinline__ Fe fe_half(const Fe& x) {
    const uint32_t mask = 0u - (x.v[0] & 1u);
    uint32_t w[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)x.v[i] + (F::p(i) & mask);
        w[i] = (uint32_t)c;
        c >>= 32;
    }
    w[8] = (uint32_t)c;
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (w[i] >> 1) | (w[i + 1] << 31);
    return r;
}
// Toom evaluation set (0, 1, -1, inf) of a cubic -> the reference's evaluation set (0, 1, 2, 3).
// v = {S(0), S(1), S(-1), c3}.  Exact field arithmetic, so the published values are the same field elements
// the direct evaluation at t = 2, 3 produces.
template <class F>
__device__ __forceinline__ void toom_to_evals(Fe* v) {
    const Fe s0 = v[0], s1 = v[1], sm = v[2], c3 = v[3];
    const Fe c2 = fe_sub<F>(fe_half<F>(fe_add<F>(s1, sm)), s0);
    const Fe c1 = fe_sub<F>(fe_half<F>(fe_sub<F>(s1, sm)), c3);
    const Fe c1x2 = fe_add<F>(c1, c1), c2x2 = fe_add<F>(c2, c2), c2x4 = fe_add<F>(c2x2, c2x2), c2x8 = fe_add<F>(c2x4, c2x4);
    const Fe c3x2 = fe_add<F>(c3, c3), c3x4 = fe_add<F>(c3x2, c3x2), c3x8 = fe_add<F>(c3x4, c3x4);
    const Fe c3x16 = fe_add<F>(c3x8, c3x8), c3x32 = fe_add<F>(c3x16, c3x16);
    // S(2) = s0 + 2 c1 + 4 c2 + 8 c3 ;  S(3) = s0 + 3 c1 + 9 c2 + 27 c3
    v[2] = fe_add<F>(fe_add<F>(s0, c1x2), fe_add<F>(c2x4, c3x8));
    const Fe c1x3 = fe_add<F>(c1x2, c1), c2x9 = fe_add<F>(c2x8, c2);
    const Fe c3x27 = fe_sub<F>(fe_sub<F>(c3x32, c3x4), c3);
    v[3] = fe_add<F>(fe_add<F>(s0, c1x3), fe_add<F>(c2x9, c3x27));
}

// What ONE thread of the last block does with the NP grid-wide sums.  finalize: derive S(1) when the launch skipped it and
// map the Toom point set back to t = 0..D (both maps are linear, so on a sharded run they are applied to the partial
// sums before the exchange).  publish: device buffer, mapped host buffer, all-reduce lanes (NCCL fallback), re-arm the
// ticket and the work counter, store the completion flag the host spins on.
template <class F, int NP, bool TOOM = false>
__device__ __forceinline__ void finalize_evals(Fe* fin, const ReduceArgs& ra) {
    if (ra.skip1 && NP > 1) fin[1] = fe_sub<F>(ra.claim, fin[0]);
    if (TOOM) toom_to_evals<F>(fin);
}
template <class F, int NP>
__device__ __forceinline__ void publish_raw(const Fe* fin, const ReduceArgs& ra, unsigned seq_override = 0) {
#pragma unroll
    for (int t = 0; t < NP; t++) {
        st_fe(ra.result_dev + ra.out_slot + t, fin[t]);
        st_fe(ra.result_host + ra.out_slot + t, fin[t]);
        if (ra.lanes) {
#pragma unroll
            for (int i = 0; i < 8; i++) ra.lanes[(ra.out_slot + t) * 8 + i] = fin[t].v[i];
        }
    }
    *ra.ticket = 0;  // ready for the next launch on this stream
    ra.ticket[kWorkCounterOffset] = 0;
    __threadfence_system();
    const unsigned seq = seq_override ? seq_override : ra.seq;
    if (seq != 0) {
        *(volatile unsigned*)ra.flag_host = seq;
        __threadfence_system();
    }
}
template <class F, int NP, bool TOOM = false>
__device__ __forceinline__ void publish_evals(Fe* fin, const ReduceArgs& ra) {
    finalize_evals<F, NP, TOOM>(fin, ra);
    publish_raw<F, NP>(fin, ra);
}

// ---- the all-reduce through the peer mailboxes (kernels.h: MailboxArgs), run by the whole last block ------------------
constexpr unsigned kMailboxTimeoutFlag = 0xffffffffu;  // published instead of the sequence number when a peer never arrives
__device__ __forceinline__ void st_fe_sys(Fe* p, const Fe& r) {  // peer (NVLink) store
    asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]),
                 "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p)
                 : "memory");
}
__device__ __forceinline__ Fe ld_fe_volatile(const Fe* p) {
    Fe r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(reinterpret_cast<const uint32_t*>(p) + 4)
                 : "memory");
    return r;
}
// s_fin[NP] (shared): in = this rank's partial sums, out = the sums over all ranks (identical on every rank).
// s_in: shared scratch of kMaxRanks * NP elements.  Returns false (on every thread) if a peer did not arrive in time.
template <class F, int NP>
__device__ __forceinline__ bool mailbox_allreduce(Fe* s_fin, Fe* s_in, unsigned* s_ok, const MailboxArgs& mb) {
    const unsigned par = mb.seq & 1u;
    if (threadIdx.x == 0) *s_ok = 1u;
    __syncthreads();
    if ((int)threadIdx.x < mb.world) {
        // thread q: my partials -> slot [par][my rank] of rank q's mailbox, then the flag
        MailboxSlot* dst = mb.peer[threadIdx.x] + par * kMaxRanks + mb.rank;
#pragma unroll 1
        for (int t = 0; t < NP; t++) st_fe_sys(dst->v + t, s_fin[t]);
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&dst->flag), "r"(mb.seq) : "memory");
        // thread q: wait for rank q's partials in my mailbox
        const MailboxSlot* src = mb.mine + par * kMaxRanks + threadIdx.x;
        unsigned f;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&src->flag) : "memory");
            if (f == mb.seq) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 4000000000ull) {  // 4 s: a peer died or never launched — fail instead of hanging the GPU
                *s_ok = 0u;
                break;
            }
        }
#pragma unroll 1
        for (int t = 0; t < NP; t++) s_in[threadIdx.x * NP + t] = ld_fe_volatile(src->v + t);
    }
    __syncthreads();
    if ((int)threadIdx.x < NP) {
        Fe v = s_in[threadIdx.x];
#pragma unroll 1
        for (int q = 1; q < mb.world; q++) v = fe_add<F>(v, s_in[q * NP + threadIdx.x]);
        s_fin[threadIdx.x] = v;
    }
    __syncthreads();
    return *s_ok != 0u;
}
// Tail of every reducing kernel, run by the WHOLE last block once s_fin[NP] (shared) holds the grid-wide sums.
template <class F, int NP, bool TOOM = false>
__device__ __forceinline__ void finish_last_block(Fe* s_fin, const ReduceArgs& ra) {
    __shared__ Fe s_in[kMaxRanks * NP];
    __shared__ unsigned s_ok;
    if (threadIdx.x == 0) {
        Fe fin[NP];
#pragma unroll
        for (int t = 0; t < NP; t++) fin[t] = s_fin[t];
        finalize_evals<F, NP, TOOM>(fin, ra);
        if (ra.mbox.world > 1) {
#pragma unroll
            for (int t = 0; t < NP; t++) s_fin[t] = fin[t];
        } else {
            publish_raw<F, NP>(fin, ra);
        }
    }
    if (ra.mbox.world > 1) {  // uniform over the block
        __syncthreads();
        const bool ok = mailbox_allreduce<F, NP>(s_fin, s_in, &s_ok, ra.mbox);
        if (threadIdx.x == 0) {
            Fe fin[NP];
#pragma unroll
            for (int t = 0; t < NP; t++) fin[t] = s_fin[t];
            publish_raw<F, NP>(fin, ra, ok ? 0u : kMailboxTimeoutFlag);
        }
    }
}

template <class F, int NP, bool TOOM = false>
__device__ __forceinline__ void reduce_publish(Fe* acc, const ReduceArgs& ra) {
    __shared__ Fe sh[NP][kWarps];
    __shared__ Fe s_fin[NP];
    __shared__ unsigned s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < NP; t++) {
        Fe v = warp_sum<F>(acc[t]);
        if (lane == 0) sh[t][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll 1
        for (int t = 0; t < NP; t++) {
            Fe v = (lane < kWarps) ? sh[t][lane] : fe_zero<F>();
            v = warp_sum<F>(v, kWarps);
            if (lane == 0) st_fe(ra.block_partials + (size_t)blockIdx.x * NP + t, v);
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned tk = atomicAdd(ra.ticket, 1u);
        s_last = (tk == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll 1
    for (int t = 0; t < NP; t++) {
        Fe v = fe_zero<F>();
        for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads)
            v = fe_add<F>(v, ld_fe_cg(ra.block_partials + (size_t)b * NP + t));
        v = warp_sum<F>(v);
        __syncthreads();  // sh reuse across t
        if (lane == 0) sh[0][warp] = v;
        __syncthreads();
        if (warp == 0) {
            Fe w = (lane < kWarps) ? sh[0][lane] : fe_zero<F>();
            w = warp_sum<F>(w, kWarps);
            if (lane == 0) s_fin[t] = w;
        }
    }
    __syncthreads();
    finish_last_block<F, NP, TOOM>(s_fin, ra);
}

// ---- launch helpers ----------------------------------------------------------------------------
template <class K>
int blocks_per_sm(K kernel, int threads) {
    int n = 0;  // a host-side query of a few microseconds: not cached (its callers are not on the round loop)
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) n = 1;
    return n;
}
inline unsigned grid_for(uint64_t items, int threads, int num_sms, int bpsm) {
    uint64_t need = (items + threads - 1) / threads;
    uint64_t cap = (uint64_t)num_sms * bpsm;
    if (cap > (uint64_t)kMaxGridBlocks) cap = kMaxGridBlocks;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

inline ReduceArgs make_ra(const ReduceScratch& s, int slot) {
    return ReduceArgs{s.block_partials, s.ticket, s.result_dev, s.result_host_devptr, slot, s.flag_host_devptr, s.seq, s.lanes, 0, Fe{}, s.mbox};
}

// host: the multiples r * 2^(32 i + 64) mod p the kernels' fe_mul_fixed consumes (r in Montgomery form)
template <class F>
FixedMul make_fixed(const Fe& r) {
    host::Field HF(F::ID);
    host::El rm;
    std::memcpy(rm.v, r.v, 32);
    FixedMul t;
    host::fixed_mul_table(HF, rm, t.v);
    return t;
}

// host-side Montgomery form of a small integer: t * R mod p by repeated addition of R (t <= kMaxDegree)
template <class F>
Fe host_small_mont(unsigned t) {
    // 8x32 limb add/sub on the host
    uint32_t acc[8] = {0};
    for (unsigned it = 0; it < t; it++) {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (uint64_t)acc[i] + F::one(i); acc[i] = (uint32_t)c; c >>= 32; }
        // conditional subtract p
        uint32_t tmp[8]; int64_t b = 0;
        for (int i = 0; i < 8; i++) { int64_t d = (int64_t)acc[i] - F::p(i) + b; tmp[i] = (uint32_t)d; b = d >> 32; }
        if (c || b == 0) for (int i = 0; i < 8; i++) acc[i] = tmp[i];
    }
    Fe r;
    for (int i = 0; i < 8; i++) r.v[i] = acc[i];
    return r;
}

// host: the FP64 multiples of the challenge for fe_fold_fixed_f64* (field_f64.cuh), both copies of the selector
template <class F>
FixedMulF64Sel make_fixed_f64(const Fe& r) {
    host::Field HF(F::ID);
    host::El rm;
    std::memcpy(rm.v, r.v, 32);
    FixedMulF64Sel t;
    host::fixed_mul_table_f64(HF, rm, t.t[0].t);
    t.t[1] = t.t[0];
    return t;
}

}  // namespace
}  // namespace zk
