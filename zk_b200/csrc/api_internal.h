// api_internal.h — what the translation units of the C ABI share (api.cu: contexts, tables, MLE / ProductPoly steps,
// transcript, field helpers; api_sumcheck.cu: prover, sum of products, verifier; api_ntt.cu: single- and multi-GPU NTT):
// the opaque handle types, the run-time NCCL symbol table, status / error helpers.  Not part of the public boundary.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/zk_b200.h"
#include "keccak.hpp"
#include "kernels.h"

using zk::Fe;
using zk::host::El;
using zk::host::Field;

// ---- NCCL, resolved at run time (only sharded contexts need it) ------------------------------------
namespace zkapi {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    // point-to-point groups: the all-to-all exchanges of the multi-GPU NTT
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    bool ok = false;
    bool p2p_ok = false;
};
constexpr int kNcclUint8 = 1, kNcclUint64 = 5, kNcclSum = 0;
NcclApi& nccl();  // resolved once, thread-safely, on first use (api.cu)
}  // namespace zkapi

struct zk_ctx {
    int device = 0;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    zk::ReduceScratch scratch{};
    zkapi::NcclComm comm = nullptr;
    uint64_t* lanes = nullptr;  // (kMaxDegree+1)*8 u64 lanes for the exact all-reduce
    // the all-reduce fused into the reducing launch: peer-mapped mailboxes (kernels.h: MailboxArgs); NCCL stays the fallback
    zk::MailboxSlot* mbox_mine = nullptr;
    zk::MailboxSlot* mbox_peer[zk::kMaxRanks] = {};
    bool mbox_ready = false;
    bool mbox_in_flight = false;  // the reduction in flight all-reduces itself (no NCCL call, no narrowing launch)
    unsigned mbox_seq = 0;
    uint64_t gather_threshold = 4096;
    std::string last_error;
    int launches = 0;
    uint64_t launches_total = 0;
    std::vector<cudaEvent_t> events;
    std::vector<float> round_ms;
    double prove_ms[3] = {0, 0, 0};
    std::vector<zk::NttPlan*> ntt_plans;  // small cache: twiddle tables are reused across calls
    unsigned cur_seq = 0;                 // sequence number of the reduction in flight
    Fe* gather_buf = nullptr;             // persistent staging for the residual all-gather (grow-only)
    size_t gather_cap = 0;                // elements
    Fe* ntt_host_buf = nullptr;           // device landing buffer of zk_ntt_host, kept between calls
    size_t ntt_host_cap = 0;              // elements
    Fe* eval_buf = nullptr;               // persistent half-size work table of zk_mle_evaluate (grow-only)
    size_t eval_cap = 0;                  // elements
    std::vector<cudaStream_t> copy_streams;  // extra H2D streams of zk_sumcheck_prove_host (lazily created)
    cudaEvent_t copy_done = nullptr;
    std::vector<cudaEvent_t> slice_events;   // zk_sumcheck_prove_host's overlapped round 0: copy-landed and kernel timing events per slice
    // device landing buffers of zk_sumcheck_prove_host, kept between calls (grow-only): a cudaMalloc + cudaFree of
    // gigabytes per proof is milliseconds of the host-buffer path and a device-wide synchronisation
    std::vector<Fe*> host_prove_buf;
    std::vector<size_t> host_prove_cap;  // elements
};

struct zk_table {
    zk_ctx* ctx;
    int field;
    unsigned n_vars;      // global number of variables
    uint64_t local_len;   // entries held by this rank
    Fe* data;
    size_t capacity;      // elements allocated
};

struct zk_transcript {
    zk::host::Transcript t;
};

namespace zkapi {

extern const char* const kMessages[];
int fail(zk_ctx* ctx, int status, const std::string& detail = std::string());
int cuda_fail(zk_ctx* ctx, cudaError_t e, const char* where);
#define CU(ctx, call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// every reducing launch publishes a fresh non-zero sequence number to the mapped completion flag
// `will_allreduce`: the launch's result is a per-rank partial — it widens it into the all-reduce lanes and
// leaves the flag alone; the narrowing kernel after the all-reduce publishes the sequence number instead.
// `single_launch`: the reduction is ONE reducing launch (every fused path); then a sharded context with mailboxes lets
// that launch all-reduce through them and publish the final result itself.
inline void next_seq(zk_ctx* ctx, bool will_allreduce = false, bool single_launch = true) {
    ++ctx->cur_seq;
    if (ctx->cur_seq == 0 || ctx->cur_seq == 0xffffffffu) ctx->cur_seq = 1;  // 0 = "do not publish", ~0 = mailbox timeout
    const bool sharded = will_allreduce && ctx->world > 1;
    const bool mbox = sharded && ctx->mbox_ready && single_launch;
    ctx->mbox_in_flight = mbox;
    ctx->scratch.seq = (sharded && !mbox) ? 0u : ctx->cur_seq;
    ctx->scratch.lanes = (sharded && !mbox) ? ctx->lanes : nullptr;
    ctx->scratch.mbox.world = 0;
    if (mbox) {
        if (++ctx->mbox_seq == 0) ctx->mbox_seq = 1;  // every rank runs the same sequence of collective reductions
        ctx->scratch.mbox.world = ctx->world;
        ctx->scratch.mbox.rank = ctx->rank;
        ctx->scratch.mbox.seq = ctx->mbox_seq;
        ctx->scratch.mbox.mine = ctx->mbox_mine;
        for (int q = 0; q < ctx->world; q++) ctx->scratch.mbox.peer[q] = ctx->mbox_peer[q];
    }
}
inline void count(zk_ctx* ctx) {
    ctx->launches_total += (uint64_t)ctx->launches;
    ctx->launches = 0;
}
inline Fe fe_from_u64x4(const uint64_t v[4]) {
    Fe r;
    std::memcpy(r.v, v, 32);
    return r;
}
inline El el_from(const uint64_t v[4]) {
    El r;
    std::memcpy(r.v, v, 32);
    return r;
}
inline unsigned log2_exact(uint64_t x) {
    unsigned l = 0;
    while (((uint64_t)1 << l) < x) l++;
    return l;
}
inline bool valid_field(int f) { return f == ZK_BLS12_381_FR || f == ZK_BLS12_377_FR; }

int table_alloc(zk_ctx* ctx, int field, unsigned n_vars, uint64_t local_len, zk_table** out);
int product_check(zk_ctx* ctx, const zk_table* const* tables, unsigned m, bool device_limits);
int distinct_check(zk_ctx* ctx, const zk_table* const* tables, unsigned m);  // consuming / in-place paths
zk::TablePtrs ptrs_of(const zk_table* const* tables, unsigned m);
// After a reducing kernel: (sharded) all-reduce the `count` partial elements exactly, then wait for the result in
// pinned host memory and copy it out.
int finish_reduction(zk_ctx* ctx, int field, int count_elems, uint64_t* out, bool allreduce);

}  // namespace zkapi
