// api.cu — the C ABI (include/zk_b200.h) over the sm_100a kernels: context/stream/scratch management,
// the Fiat-Shamir round loop (host Keccak between kernel launches), the verifier, NCCL glue for the
// sharded prover.  No CPU fallback exists: every data-path entry launches kernels on the context's GPU.
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
namespace {
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

NcclApi load_nccl() {
    NcclApi api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
    api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))dlsym(api.handle, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllGather");
    api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather;
    api.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclRecv");
    api.GroupStart = (int (*)())dlsym(api.handle, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(api.handle, "ncclGroupEnd");
    api.p2p_ok = api.ok && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    return api;
}
NcclApi& nccl() {  // resolved once, thread-safely (contexts of one process may be created from different threads)
    static NcclApi api = load_nccl();
    return api;
}
}  // namespace

struct zk_ctx {
    int device = 0;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    zk::ReduceScratch scratch{};
    NcclComm comm = nullptr;
    uint64_t* lanes = nullptr;  // (kMaxDegree+1)*8 u64 lanes for the exact all-reduce
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
    Fe* eval_buf = nullptr;               // persistent half-size work table of zk_mle_evaluate (grow-only)
    size_t eval_cap = 0;                  // elements
    std::vector<cudaStream_t> copy_streams;  // extra H2D streams of zk_sumcheck_prove_host (lazily created)
    cudaEvent_t copy_done = nullptr;
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

namespace {

const char* kMessages[] = {
    "ok",
    "evaluation vec len should equal 2^n_vars",
    "evaluate must assign to all variables",
    "cannot create product polynomial from empty polynomials",
    "cannot create product polynomial from polynomial that don't share the same number of variables",
    "invalid proof: require 1 round poly for each variable in poly",
    "couldn't evaluate initial poly",
    "verifier check failed: claimed_sum != p(0) + p(1)",
    "verification failed: initial poly evaluation != claimed sum",
    "values must be a power of 2",
    "called `Option::unwrap()` on a `None` value",
    "attempt to subtract with overflow",
    "invalid argument",
    "unsupported configuration",
    "CUDA error",
    "NCCL error",
    "out of device memory",
};

int fail(zk_ctx* ctx, int status, const std::string& detail = std::string()) {
    if (ctx) {
        ctx->last_error = kMessages[status];
        if (!detail.empty()) ctx->last_error += ": " + detail;
    }
    return status;
}
int cuda_fail(zk_ctx* ctx, cudaError_t e, const char* where) {
    int st = (e == cudaErrorMemoryAllocation) ? ZK_ERR_OOM : ZK_ERR_CUDA;
    return fail(ctx, st, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// every reducing launch publishes a fresh non-zero sequence number to the mapped completion flag
// `will_allreduce`: the launch's result is a per-rank partial — it widens it into the all-reduce lanes and
// leaves the flag alone; the narrowing kernel after the all-reduce publishes the sequence number instead.
inline void next_seq(zk_ctx* ctx, bool will_allreduce = false) {
    if (++ctx->cur_seq == 0) ctx->cur_seq = 1;
    const bool sharded = will_allreduce && ctx->world > 1;
    ctx->scratch.seq = sharded ? 0u : ctx->cur_seq;
    ctx->scratch.lanes = sharded ? ctx->lanes : nullptr;
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
inline bool valid_field(int f) { return f == ZK_BLS12_381_FR || f == ZK_BLS12_377_FR; }
inline unsigned log2_exact(uint64_t x) {
    unsigned l = 0;
    while (((uint64_t)1 << l) < x) l++;
    return l;
}

int table_alloc(zk_ctx* ctx, int field, unsigned n_vars, uint64_t local_len, zk_table** out) {
    zk_table* t = new (std::nothrow) zk_table{ctx, field, n_vars, local_len, nullptr, (size_t)local_len};
    if (!t) return fail(ctx, ZK_ERR_OOM);
    cudaError_t e = cudaMalloc((void**)&t->data, (size_t)(local_len ? local_len : 1) * sizeof(Fe));
    if (e != cudaSuccess) {
        delete t;
        cudaGetLastError();
        return cuda_fail(ctx, e, "cudaMalloc(table)");
    }
    *out = t;
    return ZK_OK;
}

int product_check(zk_ctx* ctx, const zk_table* const* tables, unsigned m, bool device_limits) {
    if (m == 0 || tables == nullptr) return fail(ctx, ZK_ERR_EMPTY_PRODUCT);
    for (unsigned k = 0; k < m; k++)
        if (!tables[k]) return fail(ctx, ZK_ERR_INVALID_ARG, "null table");
    for (unsigned k = 1; k < m; k++)
        if (tables[k]->n_vars != tables[0]->n_vars) return fail(ctx, ZK_ERR_NVARS_MISMATCH);
    for (unsigned k = 1; k < m; k++)
        if (tables[k]->field != tables[0]->field || tables[k]->local_len != tables[0]->local_len)
            return fail(ctx, ZK_ERR_INVALID_ARG, "factors must share field and sharding");
    if (device_limits && m > ZK_MAX_FACTORS) return fail(ctx, ZK_ERR_UNSUPPORTED, "more than ZK_MAX_FACTORS factors");
    return ZK_OK;
}
zk::TablePtrs ptrs_of(const zk_table* const* tables, unsigned m) {
    zk::TablePtrs p{};
    for (unsigned k = 0; k < m && k < (unsigned)zk::kMaxFactors; k++) p.t[k] = tables[k]->data;
    return p;
}

// After a reducing kernel: (sharded) all-reduce the `count` partial elements exactly, then wait for the
// result in pinned host memory and copy it out.
int finish_reduction(zk_ctx* ctx, int field, int count_elems, uint64_t* out, bool allreduce) {
    if (allreduce && ctx->world > 1) {
        // the reducing launch already wrote one 32-bit limb per u64 lane (ReduceScratch::lanes)
        int rc = nccl().AllReduce(ctx->lanes, ctx->lanes, (size_t)count_elems * 8, kNcclUint64, kNcclSum, ctx->comm,
                                  ctx->stream);
        if (rc != 0) return fail(ctx, ZK_ERR_NCCL, nccl().GetErrorString ? nccl().GetErrorString(rc) : "allreduce");
        CU(ctx, zk::launch_narrow(field, ctx->lanes, ctx->scratch.result_dev, ctx->scratch.result_host_devptr, count_elems,
                                  ctx->scratch.flag_host_devptr, ctx->cur_seq, ctx->stream, &ctx->launches));
    }
    // spin on the mapped completion flag the last block (or the narrowing kernel) stores after the results: a few
    // microseconds cheaper per round than a stream synchronisation; the stream status is consulted every so
    // often so that a failed launch cannot hang the caller
    volatile unsigned* flag = ctx->scratch.flag_host;
    const unsigned want = ctx->cur_seq;
    for (unsigned spins = 0; *flag != want; spins++) {
        if ((spins & 0x3fff) == 0x3fff) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) {
                if (*flag != want) CU(ctx, cudaStreamSynchronize(ctx->stream));
                break;
            }
            if (q != cudaErrorNotReady) return cuda_fail(ctx, q, "round kernel");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (out) std::memcpy(out, ctx->scratch.result_host, (size_t)count_elems * 32);
    return ZK_OK;
}

int ctx_init(zk_ctx* c) {
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    CU(c, cudaGetDeviceProperties(&prop, c->device));
    c->scratch.num_sms = prop.multiProcessorCount;
    const size_t np = (size_t)zk::kMaxGridBlocks * (zk::kMaxDegree + 1);
    CU(c, cudaMalloc((void**)&c->scratch.block_partials, np * sizeof(Fe)));
    // ticket word + (128 bytes further) the 64-bit work counter of the round kernels
    CU(c, cudaMalloc((void**)&c->scratch.ticket, 256));
    CU(c, cudaMemset(c->scratch.ticket, 0, 256));
    CU(c, cudaMalloc((void**)&c->scratch.result_dev, (zk::kMaxDegree + 1) * sizeof(Fe)));
    CU(c, cudaHostAlloc((void**)&c->scratch.result_host, (zk::kMaxDegree + 1) * sizeof(Fe), cudaHostAllocMapped));
    CU(c, cudaHostGetDevicePointer((void**)&c->scratch.result_host_devptr, c->scratch.result_host, 0));
    CU(c, cudaHostAlloc((void**)&c->scratch.flag_host, 64, cudaHostAllocMapped));
    *c->scratch.flag_host = 0;
    CU(c, cudaHostGetDevicePointer((void**)&c->scratch.flag_host_devptr, c->scratch.flag_host, 0));
    c->scratch.seq = 0;
    CU(c, cudaMalloc((void**)&c->lanes, (zk::kMaxDegree + 1) * 8 * sizeof(uint64_t)));
    c->events.resize(2 * 260);
    for (auto& ev : c->events) CU(c, cudaEventCreate(&ev));
    return ZK_OK;
}

}  // namespace

extern "C" {

const char* zk_status_string(int status) {
    if (status < 0 || status > ZK_ERR_OOM) return "unknown status";
    return kMessages[status];
}
const char* zk_last_error(const zk_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int zk_ctx_create(int device, zk_ctx** out) {
    if (!out) return ZK_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return ZK_ERR_CUDA;  // no GPU: there is no CPU fallback
    }
    if (device < 0 || device >= ndev) return ZK_ERR_INVALID_ARG;
    zk_ctx* c = new (std::nothrow) zk_ctx();
    if (!c) return ZK_ERR_OOM;
    c->device = device;
    int st = ctx_init(c);
    if (st != ZK_OK) {
        std::fprintf(stderr, "zk_ctx_create: %s\n", c->last_error.c_str());
        delete c;
        return st;
    }
    *out = c;
    return ZK_OK;
}

int zk_nccl_unique_id(void* id_out_128) {
    if (!id_out_128) return ZK_ERR_INVALID_ARG;
    if (!nccl().ok) return ZK_ERR_NCCL;
    NcclUniqueId id;
    if (nccl().GetUniqueId(&id) != 0) return ZK_ERR_NCCL;
    std::memcpy(id_out_128, &id, 128);
    return ZK_OK;
}

int zk_ctx_create_sharded(int device, int rank, int world, const void* nccl_id, zk_ctx** out) {
    if (!out || world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) return ZK_ERR_INVALID_ARG;
    int st = zk_ctx_create(device, out);
    if (st != ZK_OK) return st;
    zk_ctx* c = *out;
    c->rank = rank;
    c->world = world;
    if (world > 1) {
        if (!nccl_id || !nccl().ok) {
            zk_ctx_destroy(c);
            *out = nullptr;
            return ZK_ERR_NCCL;
        }
        NcclUniqueId id;
        std::memcpy(&id, nccl_id, 128);
        int rc = nccl().CommInitRank(&c->comm, world, id, rank);
        if (rc != 0) {
            std::fprintf(stderr, "ncclCommInitRank: %s\n", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
            zk_ctx_destroy(c);
            *out = nullptr;
            return ZK_ERR_NCCL;
        }
    }
    return ZK_OK;
}

void zk_ctx_destroy(zk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    for (auto* pl : c->ntt_plans) zk::ntt_plan_destroy(pl);
    cudaFree(c->gather_buf);
    cudaFree(c->eval_buf);
    for (auto ev : c->events) cudaEventDestroy(ev);
    cudaFree(c->scratch.block_partials);
    cudaFree(c->scratch.ticket);
    cudaFree(c->scratch.result_dev);
    cudaFreeHost(c->scratch.result_host);
    cudaFreeHost(c->scratch.flag_host);
    cudaFree(c->lanes);
    for (Fe* b : c->host_prove_buf) cudaFree(b);
    for (cudaStream_t s : c->copy_streams) cudaStreamDestroy(s);
    if (c->copy_done) cudaEventDestroy(c->copy_done);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
int zk_ctx_rank(const zk_ctx* c) { return c ? c->rank : -1; }
int zk_ctx_world(const zk_ctx* c) { return c ? c->world : -1; }
int zk_ctx_set_gather_threshold(zk_ctx* c, uint64_t local_len) {
    if (!c || local_len < 1) return ZK_ERR_INVALID_ARG;
    c->gather_threshold = local_len;
    return ZK_OK;
}
uint64_t zk_ctx_launch_count(const zk_ctx* c) { return c ? c->launches_total + (uint64_t)c->launches : 0; }
unsigned zk_ctx_last_round_ms(const zk_ctx* c, float* ms_out, unsigned cap) {
    if (!c || !ms_out) return 0;
    unsigned n = (unsigned)c->round_ms.size();
    if (n > cap) n = cap;
    for (unsigned i = 0; i < n; i++) ms_out[i] = c->round_ms[i];
    return n;
}
int zk_ctx_last_prove_ms(const zk_ctx* c, double out[3]) {
    if (!c || !out) return ZK_ERR_INVALID_ARG;
    for (int i = 0; i < 3; i++) out[i] = c->prove_ms[i];
    return ZK_OK;
}
int zk_ctx_synchronize(zk_ctx* c) {
    if (!c) return ZK_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    return ZK_OK;
}

void* zk_ctx_stream(const zk_ctx* c) { return c ? (void*)c->stream : nullptr; }

int zk_host_alloc(size_t bytes, void** out) {
    if (!out) return ZK_ERR_INVALID_ARG;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ZK_ERR_OOM;
    }
    return ZK_OK;
}
int zk_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return ZK_ERR_CUDA;
    }
    return ZK_OK;
}

// ---- tables ------------------------------------------------------------------------------------
int zk_table_upload(zk_ctx* ctx, int field, const uint64_t* mont_aos, uint64_t len, unsigned n_vars, zk_table** out) {
    if (!ctx || !out || !valid_field(field)) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (n_vars >= 48 || len != ((uint64_t)1 << n_vars)) return fail(ctx, ZK_ERR_EVAL_LEN);
    if (!mont_aos) return fail(ctx, ZK_ERR_INVALID_ARG, "null evaluations");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t world = (uint64_t)ctx->world;
    if (world > 1 && len < world) return fail(ctx, ZK_ERR_UNSUPPORTED, "table smaller than the number of ranks");
    const uint64_t local = len / world;
    int st = table_alloc(ctx, field, n_vars, local, out);
    if (st != ZK_OK) return st;
    if (world == 1) {
        CU(ctx, cudaMemcpyAsync((*out)->data, mont_aos, (size_t)len * 32, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // rank q keeps global[j*world + q]: a strided 2-D copy (32-byte rows, pitch world*32)
        CU(ctx, cudaMemcpy2DAsync((*out)->data, 32, mont_aos + 4 * (uint64_t)ctx->rank, (size_t)world * 32, 32,
                                  (size_t)local, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_table_generate(zk_ctx* ctx, int field, uint64_t seed, uint64_t table_id, unsigned n_vars, zk_table** out) {
    if (!ctx || !out || !valid_field(field) || n_vars >= 40) return fail(ctx, ZK_ERR_INVALID_ARG);
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t len = (uint64_t)1 << n_vars, world = (uint64_t)ctx->world;
    if (world > 1 && len < world) return fail(ctx, ZK_ERR_UNSUPPORTED, "table smaller than the number of ranks");
    const uint64_t local = len / world;
    int st = table_alloc(ctx, field, n_vars, local, out);
    if (st != ZK_OK) return st;
    CU(ctx, zk::launch_generate(field, (*out)->data, local, seed, table_id, (uint64_t)ctx->rank, world, ctx->stream,
                                &ctx->launches));
    count(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_table_regenerate(zk_ctx* ctx, zk_table* t, uint64_t seed, uint64_t table_id) {
    if (!ctx || !t) return fail(ctx, ZK_ERR_INVALID_ARG);
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t world = (uint64_t)ctx->world, local = ((uint64_t)1 << t->n_vars) / world;
    if (local > t->capacity) return fail(ctx, ZK_ERR_INVALID_ARG, "table allocation too small");
    t->local_len = local;
    CU(ctx, zk::launch_generate(t->field, t->data, local, seed, table_id, (uint64_t)ctx->rank, world, ctx->stream,
                                &ctx->launches));
    count(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_table_clone(zk_ctx* ctx, const zk_table* in, zk_table** out) {
    if (!ctx || !in || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    CU(ctx, cudaSetDevice(ctx->device));
    int st = table_alloc(ctx, in->field, in->n_vars, in->local_len, out);
    if (st != ZK_OK) return st;
    CU(ctx, cudaMemcpyAsync((*out)->data, in->data, (size_t)in->local_len * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
void zk_table_free(zk_table* t) {
    if (!t) return;
    if (t->ctx) cudaSetDevice(t->ctx->device);
    cudaFree(t->data);
    delete t;
}
unsigned zk_table_n_vars(const zk_table* t) { return t ? t->n_vars : 0; }
uint64_t zk_table_local_len(const zk_table* t) { return t ? t->local_len : 0; }
int zk_table_field(const zk_table* t) { return t ? t->field : -1; }

int zk_table_download(zk_ctx* ctx, const zk_table* t, uint64_t* out) {
    if (!ctx || !t || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(out, t->data, (size_t)t->local_len * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// ---- MLE ops ------------------------------------------------------------------------------------
int zk_mle_partial_evaluate(zk_ctx* ctx, const zk_table* in, unsigned initial_var, const uint64_t* assignments,
                            unsigned n_assign, zk_table** out) {
    if (!ctx || !in || !out || (n_assign && !assignments)) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (ctx->world > 1) return fail(ctx, ZK_ERR_UNSUPPORTED, "partial_evaluate on a sharded context");
    CU(ctx, cudaSetDevice(ctx->device));
    // Rust: index_pair((n_vars - i) as u8, initial_var as u8) underflows when a step has no such variable
    if ((uint64_t)initial_var + n_assign > in->n_vars) return fail(ctx, ZK_ERR_VAR_RANGE);
    const unsigned n = in->n_vars;
    if (n_assign == 0) return zk_table_clone(ctx, in, out);
    // ping-pong between two half-size buffers; the first step reads the input table directly
    zk_table *a = nullptr, *b = nullptr;
    int st = table_alloc(ctx, in->field, n - 1, in->local_len / 2, &a);
    if (st != ZK_OK) return st;
    if (n_assign > 1) {
        st = table_alloc(ctx, in->field, n - 2, in->local_len / 4, &b);
        if (st != ZK_OK) {
            zk_table_free(a);
            return st;
        }
    }
    const Fe* src = in->data;
    zk_table* dst = a;
    for (unsigned s = 0; s < n_assign; s++) {
        cudaError_t e = zk::launch_fold_var(in->field, src, dst->data, n - s, initial_var,
                                            fe_from_u64x4(assignments + 4 * (size_t)s), ctx->stream, &ctx->launches);
        if (e != cudaSuccess) {
            zk_table_free(a);
            zk_table_free(b);
            return cuda_fail(ctx, e, "fold_var");
        }
        dst->n_vars = n - s - 1;
        dst->local_len = (uint64_t)1 << dst->n_vars;
        src = dst->data;
        dst = (dst == a) ? b : a;
    }
    count(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    zk_table* result = (dst == a) ? b : a;  // the last written
    if (n_assign == 1) result = a;
    zk_table_free(result == a ? b : a);
    *out = result;
    return ZK_OK;
}

int zk_mle_evaluate(zk_ctx* ctx, const zk_table* in, const uint64_t* point, unsigned len, uint64_t out[4]) {
    if (!ctx || !in || !out || (len && !point)) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (len != in->n_vars) return fail(ctx, ZK_ERR_EVALUATE_ARITY);
    if (ctx->world > 1) return fail(ctx, ZK_ERR_UNSUPPORTED, "evaluate on a sharded context");
    CU(ctx, cudaSetDevice(ctx->device));
    if (len == 0) {
        CU(ctx, cudaMemcpyAsync(out, in->data, 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return ZK_OK;
    }
    // first fold out of place into a half-size work buffer, the rest in place (the reference folds inside its
    // private clone, evaluation_form.rs:49-72).  The buffer is kept by the context: a cudaMalloc/cudaFree pair
    // costs more than all the folds of a 2^20-entry table.  All launches are queued back to back (the
    // assignments are known up front); one synchronisation at the end.
    const uint64_t half = in->local_len / 2;
    if (ctx->eval_cap < half) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->eval_buf);
        ctx->eval_buf = nullptr;
        ctx->eval_cap = 0;
        CU(ctx, cudaMalloc((void**)&ctx->eval_buf, (size_t)half * sizeof(Fe)));
        ctx->eval_cap = half;
    }
    Fe* w = ctx->eval_buf;
    cudaError_t e = zk::launch_fold_var(in->field, in->data, w, in->n_vars, 0, fe_from_u64x4(point), ctx->stream,
                                        &ctx->launches);
    zk::TablePtrs p{};
    p.t[0] = w;
    uint64_t cur = half;
    for (unsigned s = 1; s < len && e == cudaSuccess; s++) {
        e = zk::launch_fold(in->field, p, 1, cur / 2, fe_from_u64x4(point + 4 * (size_t)s), ctx->stream, &ctx->launches);
        cur /= 2;
    }
    count(ctx);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, w, 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (ctx->eval_cap * sizeof(Fe) > ((size_t)1 << 30)) {  // do not hoard more than 1 GiB between calls
        cudaFree(ctx->eval_buf);
        ctx->eval_buf = nullptr;
        ctx->eval_cap = 0;
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "evaluate");
    return ZK_OK;
}

int zk_mle_to_bytes(zk_ctx* ctx, const zk_table* in, uint8_t* out) {
    if (!ctx || !in || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t chunk = (uint64_t)1 << 20;
    uint8_t* dbuf = nullptr;
    CU(ctx, cudaMalloc((void**)&dbuf, (size_t)(in->local_len < chunk ? (in->local_len ? in->local_len : 1) : chunk) * 32));
    for (uint64_t off = 0; off < in->local_len; off += chunk) {
        uint64_t n = in->local_len - off < chunk ? in->local_len - off : chunk;
        cudaError_t e = zk::launch_to_bytes(in->field, in->data + off, n, dbuf, ctx->stream, &ctx->launches);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + off * 32, dbuf, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(dbuf);
            return cuda_fail(ctx, e, "to_bytes");
        }
    }
    count(ctx);
    cudaFree(dbuf);
    return ZK_OK;
}

// ---- ProductPoly ops ------------------------------------------------------------------------------
int zk_product_check(const zk_table* const* tables, unsigned m) { return product_check(nullptr, tables, m, false); }

int zk_product_evaluate(zk_ctx* ctx, const zk_table* const* tables, unsigned m, const uint64_t* point, unsigned len,
                        uint64_t out[4]) {
    if (!ctx || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, false);
    if (st != ZK_OK) return st;
    if (len != tables[0]->n_vars) return fail(ctx, ZK_ERR_EVALUATE_ARITY);
    Field F(tables[0]->field);
    El prod = F.one();  // product_poly.rs:41 try_fold(F::one(), ..)
    for (unsigned k = 0; k < m; k++) {
        uint64_t v[4];
        st = zk_mle_evaluate(ctx, tables[k], point, len, v);
        if (st != ZK_OK) return st;
        prod = F.mul(prod, el_from(v));
    }
    std::memcpy(out, prod.v, 32);
    return ZK_OK;
}

int zk_product_prod_reduce(zk_ctx* ctx, const zk_table* const* tables, unsigned m, zk_table** out) {
    if (!ctx || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st != ZK_OK) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    st = table_alloc(ctx, tables[0]->field, tables[0]->n_vars, tables[0]->local_len, out);
    if (st != ZK_OK) return st;
    CU(ctx, zk::launch_prod_reduce(tables[0]->field, ptrs_of(tables, m), (int)m, tables[0]->local_len, (*out)->data,
                                   ctx->stream, &ctx->launches));
    count(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_product_sum(zk_ctx* ctx, const zk_table* const* tables, unsigned m, uint64_t out[4]) {
    if (!ctx || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st != ZK_OK) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    next_seq(ctx, true);
    CU(ctx, zk::launch_product_sum(tables[0]->field, ptrs_of(tables, m), (int)m, tables[0]->local_len, ctx->scratch,
                                   ctx->stream, &ctx->launches));
    st = finish_reduction(ctx, tables[0]->field, 1, out, true);
    count(ctx);
    return st;
}

int zk_product_round_poly(zk_ctx* ctx, const zk_table* const* tables, unsigned m, unsigned degree, uint64_t* out) {
    if (!ctx || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st != ZK_OK) return st;
    if (degree > ZK_MAX_DEGREE) return fail(ctx, ZK_ERR_UNSUPPORTED, "degree > ZK_MAX_DEGREE");
    if (tables[0]->n_vars == 0 || tables[0]->local_len < 2) return fail(ctx, ZK_ERR_VAR_RANGE);
    CU(ctx, cudaSetDevice(ctx->device));
    next_seq(ctx, true);
    CU(ctx, zk::launch_round_poly(tables[0]->field, ptrs_of(tables, m), (int)m, (int)degree, tables[0]->local_len / 2,
                                  ctx->scratch, ctx->stream, &ctx->launches));
    st = finish_reduction(ctx, tables[0]->field, (int)degree + 1, out, true);
    count(ctx);
    return st;
}

int zk_product_fold_inplace(zk_ctx* ctx, zk_table* const* tables, unsigned m, const uint64_t r[4]) {
    if (!ctx || !r) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st != ZK_OK) return st;
    if (tables[0]->n_vars == 0 || tables[0]->local_len < 2) return fail(ctx, ZK_ERR_VAR_RANGE);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, zk::launch_fold(tables[0]->field, ptrs_of(tables, m), (int)m, tables[0]->local_len / 2, fe_from_u64x4(r),
                            ctx->stream, &ctx->launches));
    count(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (unsigned k = 0; k < m; k++) {
        tables[k]->n_vars -= 1;
        tables[k]->local_len /= 2;
    }
    return ZK_OK;
}

int zk_product_fold_then_round_poly(zk_ctx* ctx, zk_table* const* tables, unsigned m, unsigned degree,
                                    const uint64_t r[4], uint64_t* out) {
    if (!ctx || !r || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st != ZK_OK) return st;
    if (degree > ZK_MAX_DEGREE) return fail(ctx, ZK_ERR_UNSUPPORTED, "degree > ZK_MAX_DEGREE");
    if (tables[0]->n_vars < 2 || tables[0]->local_len < 4) return fail(ctx, ZK_ERR_VAR_RANGE);
    CU(ctx, cudaSetDevice(ctx->device));
    next_seq(ctx, true);
    CU(ctx, zk::launch_fold_round_poly(tables[0]->field, ptrs_of(tables, m), (int)m, (int)degree, tables[0]->local_len,
                                       fe_from_u64x4(r), ctx->scratch, ctx->stream, &ctx->launches));
    st = finish_reduction(ctx, tables[0]->field, (int)degree + 1, out, true);
    count(ctx);
    if (st != ZK_OK) return st;
    for (unsigned k = 0; k < m; k++) {
        tables[k]->n_vars -= 1;
        tables[k]->local_len /= 2;
    }
    return ZK_OK;
}

// ---- sumcheck prover -------------------------------------------------------------------------------
namespace {

// poly.to_bytes() absorbed into the transcript (prover.rs:16-17, verifier.rs:21-22): the device
// canonicalises + byte-swaps, the host hashes; factor-major, index-minor (product_poly.rs:77-83).
int absorb_tables(zk_ctx* ctx, const zk_table* const* tables, unsigned m, zk::host::Transcript& tr) {
    if (ctx->world > 1) return fail(ctx, ZK_ERR_UNSUPPORTED, "prove()/verify() with the initial-poly absorb on a sharded context");
    const uint64_t chunk = (uint64_t)1 << 20;  // 32 MiB per chunk, double buffered
    uint8_t* dbuf[2] = {nullptr, nullptr};
    uint8_t* hbuf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaMalloc((void**)&dbuf[i], chunk * 32);
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&hbuf[i], chunk * 32, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    }
    int st = ZK_OK;
    if (e == cudaSuccess) {
        // pipeline: while the host hashes chunk c, the device produces chunk c+1
        std::vector<std::pair<const Fe*, uint64_t>> jobs;
        for (unsigned k = 0; k < m; k++)
            for (uint64_t off = 0; off < tables[k]->local_len; off += chunk)
                jobs.push_back({tables[k]->data + off, tables[k]->local_len - off < chunk ? tables[k]->local_len - off : chunk});
        auto issue = [&](size_t c) -> cudaError_t {
            int b = (int)(c & 1);
            cudaError_t ee = zk::launch_to_bytes(tables[0]->field, jobs[c].first, jobs[c].second, dbuf[b], ctx->stream, &ctx->launches);
            if (ee == cudaSuccess) ee = cudaMemcpyAsync(hbuf[b], dbuf[b], (size_t)jobs[c].second * 32, cudaMemcpyDeviceToHost, ctx->stream);
            if (ee == cudaSuccess) ee = cudaEventRecord(done[b], ctx->stream);
            return ee;
        };
        if (!jobs.empty()) e = issue(0);
        for (size_t c = 0; c < jobs.size() && e == cudaSuccess; c++) {
            if (c + 1 < jobs.size()) e = issue(c + 1);
            if (e == cudaSuccess) e = cudaEventSynchronize(done[c & 1]);
            if (e == cudaSuccess) tr.append(hbuf[c & 1], (size_t)jobs[c].second * 32);
        }
    }
    if (e != cudaSuccess) st = cuda_fail(ctx, e, "absorb");
    for (int i = 0; i < 2; i++) {
        if (dbuf[i]) cudaFree(dbuf[i]);
        if (hbuf[i]) cudaFreeHost(hbuf[i]);
        if (done[i]) cudaEventDestroy(done[i]);  // (found by the host-mock run under AddressSanitizer: two events leaked per call)
    }
    return st;
}

// Env-gated per-phase log in the spirit of the reference's `stat` crate (stat/src/lib.rs:12-30, PERF_LOG=true).
bool perf_log_enabled() {  // a magic static: contexts may live on different threads (one thread per zk_ctx)
    static const bool on = [] {
        const char* e = std::getenv("PERF_LOG");
        return e && std::strcmp(e, "true") == 0;
    }();
    return on;
}

// Value at x of the polynomial of degree < np given by its evaluations ys[t] at t = 0..np-1 (barycentric form, no
// division by anything that depends on x): the next round's S(0) + S(1), `claimed_sum = p.evaluate(challenge)` in the
// verifier (sumcheck/src/verifier.rs:68-70).  The inverse denominators 1 / prod_{u != t} (t - u) are cached per field.
class RoundPolyEvaluator {
   public:
    RoundPolyEvaluator(const Field& F, int np) : F_(F), np_(np), w_((size_t)np), pt_((size_t)np) {
        for (int t = 0; t < np; t++) pt_[(size_t)t] = F.from_u64((uint64_t)t);
        // the inversions (a field exponentiation each) are done once per thread, field and degree
        static thread_local std::vector<El> cache[2][ZK_MAX_DEGREE + 2];
        std::vector<El>& c = cache[F.id()][np];
        if (c.empty()) {
            for (int t = 0; t < np; t++) {
                El den = F.one();
                for (int u = 0; u < np; u++)
                    if (u != t) den = F.mul(den, F.sub(pt_[(size_t)t], pt_[(size_t)u]));
                c.push_back(F.inverse(den));
            }
        }
        w_ = c;
    }
    El at(const uint64_t* ys_mont, const El& x) const {
        std::vector<El> d((size_t)np_), pre((size_t)np_ + 1), suf((size_t)np_ + 1);
        for (int u = 0; u < np_; u++) d[(size_t)u] = F_.sub(x, pt_[(size_t)u]);
        pre[0] = F_.one();
        for (int u = 0; u < np_; u++) pre[(size_t)u + 1] = F_.mul(pre[(size_t)u], d[(size_t)u]);
        suf[(size_t)np_] = F_.one();
        for (int u = np_ - 1; u >= 0; u--) suf[(size_t)u] = F_.mul(suf[(size_t)u + 1], d[(size_t)u]);
        El acc = F_.zero();
        for (int t = 0; t < np_; t++) {
            El y;
            std::memcpy(y.v, ys_mont + 4 * (size_t)t, 32);
            acc = F_.add(acc, F_.mul(F_.mul(y, w_[(size_t)t]), F_.mul(pre[(size_t)t], suf[(size_t)t + 1])));
        }
        return acc;
    }

   private:
    const Field& F_;
    int np_;
    std::vector<El> w_, pt_;
};

struct ProveTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

}  // namespace

// The round loop of prover.rs:33-73.  `sop` == nullptr: the reference's ProductPoly (the m tables are the factors);
// otherwise the polynomial is the sum of products `*sop` over the m tables — same transcript
// protocol, same folds, only the round-sum kernels differ.
static int prove_core(zk_ctx* ctx, zk_table* const* tables, unsigned m, unsigned degree, const uint64_t sum[4],
                      int absorb_initial_poly, uint64_t* round_polys_out, uint64_t* challenges_out,
                      uint64_t* final_evals_out, const zk::SopSpec* sop) {
    if (!ctx || !sum) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st != ZK_OK) return st;
    if (degree > ZK_MAX_DEGREE) return fail(ctx, ZK_ERR_UNSUPPORTED, "degree > ZK_MAX_DEGREE");
    const unsigned n = tables[0]->n_vars;
    if (n && !round_polys_out) return fail(ctx, ZK_ERR_INVALID_ARG, "null round_polys_out");
    if (n > 255) return fail(ctx, ZK_ERR_UNSUPPORTED);
    CU(ctx, cudaSetDevice(ctx->device));
    const int field = tables[0]->field;
    const Field F(field);
    const int np = (int)degree + 1;
    ProveTimer timer;
    ctx->round_ms.clear();
    ctx->prove_ms[0] = ctx->prove_ms[1] = ctx->prove_ms[2] = 0;

    zk::host::Transcript tr;  // Transcript::new()
    if (absorb_initial_poly) {
        st = absorb_tables(ctx, tables, m, tr);
        if (st != ZK_OK) return st;
        ctx->prove_ms[1] = timer.ms();
    }
    tr.append_element(F, el_from(sum));  // prover.rs:42

    // working views of the tables (they are consumed)
    zk::TablePtrs cur = ptrs_of(tables, m);
    uint64_t cur_len = tables[0]->local_len;
    bool sharded = ctx->world > 1;

    // Gather the per-rank residual tables (local length L) into full tables of L*world entries on every rank.
    auto gather = [&]() -> int {
        const uint64_t L = cur_len, G = (uint64_t)ctx->world;
        // persistent staging (grow-only): [m] all-gather landing zones + [m] interleaved tables
        const size_t need = (size_t)(2 * m * L * G);
        if (ctx->gather_cap < need) {
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->gather_buf);
            ctx->gather_buf = nullptr;
            ctx->gather_cap = 0;
            CU(ctx, cudaMalloc((void**)&ctx->gather_buf, need * sizeof(Fe)));
            ctx->gather_cap = need;
        }
        for (unsigned k = 0; k < m; k++) {
            Fe* stage = ctx->gather_buf + (size_t)k * L * G;
            Fe* full = ctx->gather_buf + (size_t)(m + k) * L * G;
            int rc = nccl().AllGather(cur.t[k], stage, (size_t)L * 32, kNcclUint8, ctx->comm, ctx->stream);
            if (rc != 0) return fail(ctx, ZK_ERR_NCCL, "allgather");
            cudaError_t e = zk::launch_interleave(stage, full, L, (unsigned)G, ctx->stream, &ctx->launches);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "gather");
            cur.t[k] = full;
        }
        cur_len = L * G;
        sharded = false;
        return ZK_OK;
    };

    // round sums of the current tables / fold at r fused with the next round's sums
    auto launch_sums = [&]() -> cudaError_t {
        return sop ? zk::launch_sop_round_poly(field, cur, *sop, (int)degree, cur_len / 2, ctx->scratch, ctx->stream, &ctx->launches)
                   : zk::launch_round_poly(field, cur, (int)m, (int)degree, cur_len / 2, ctx->scratch, ctx->stream, &ctx->launches);
    };
    auto launch_fold_sums = [&](const Fe& rf, const Fe* claim_ptr) -> cudaError_t {
        return sop ? zk::launch_sop_fold_round_poly(field, cur, *sop, (int)degree, cur_len, rf, ctx->scratch, ctx->stream, &ctx->launches, claim_ptr)
                   : zk::launch_fold_round_poly(field, cur, (int)m, (int)degree, cur_len, rf, ctx->scratch, ctx->stream, &ctx->launches, claim_ptr);
    };

    std::vector<uint64_t> S((size_t)np * 4);
    const RoundPolyEvaluator round_eval(F, np);
    size_t ev = 0;
    auto timed = [&](auto&& launch) -> cudaError_t {
        cudaError_t e = cudaEventRecord(ctx->events[ev], ctx->stream);
        if (e == cudaSuccess) e = launch();
        if (e == cudaSuccess) e = cudaEventRecord(ctx->events[ev + 1], ctx->stream);
        ev += 2;
        return e;
    };

    if (n > 0) {
        if (sharded && (cur_len < 2 || cur_len <= ctx->gather_threshold)) {
            st = gather();
            if (st != ZK_OK) return st;
        }
        cudaError_t e = timed([&] { next_seq(ctx, sharded); return launch_sums(); });
        if (e != cudaSuccess) return cuda_fail(ctx, e, "round_poly");
        if (perf_log_enabled()) {
            cudaStreamSynchronize(ctx->stream);
            std::fprintf(stderr, "[zk_b200 rank %d] round 0 kernel done at %.3f ms\n", ctx->rank, timer.ms());
        }
        st = finish_reduction(ctx, field, np, S.data(), sharded);
        if (st != ZK_OK) return st;
        if (perf_log_enabled()) std::fprintf(stderr, "[zk_b200 rank %d] round 0 reduced at %.3f ms\n", ctx->rank, timer.ms());
    }
    El r = F.zero();
    double t_prev = timer.ms();
    for (unsigned round = 0; round < n; round++) {
        if (perf_log_enabled()) {
            double t_now = timer.ms();
            std::fprintf(stderr, "[zk_b200 rank %d] round %u: table 2^%u%s, %.3f ms since previous round\n", ctx->rank, round,
                         log2_exact(cur_len), sharded ? " (sharded)" : "", t_now - t_prev);
            t_prev = t_now;
        }
        std::memcpy(round_polys_out + (size_t)round * np * 4, S.data(), (size_t)np * 32);
        for (int t = 0; t < np; t++) tr.append_element(F, el_from(S.data() + 4 * t));  // prover.rs:59
        r = tr.sample_field_element(F);                                                // prover.rs:62
        if (challenges_out) std::memcpy(challenges_out + 4 * (size_t)round, r.v, 32);
        if (round + 1 == n) break;
        const Fe rf = fe_from_u64x4(r.v);
        cudaError_t e;
        if (sharded && cur_len / 2 <= ctx->gather_threshold) {
            // fold locally, gather the residual, continue unsharded
            e = timed([&] {
                cudaError_t ee = zk::launch_fold(field, cur, (int)m, cur_len / 2, rf, ctx->stream, &ctx->launches);
                cur_len /= 2;
                return ee;
            });
            if (e != cudaSuccess) return cuda_fail(ctx, e, "fold");
            st = gather();
            if (st != ZK_OK) return st;
            e = timed([&] { next_seq(ctx, sharded); return launch_sums(); });
        } else {
            // S_{round+1}(0) + S_{round+1}(1) = S_round(r): the kernel skips the t = 1 term and derives it (sharded:
            // the map is linear, so the value goes to rank 0 and zero to the others before the all-reduce).
            // Only when the D+1 evaluations determine the round polynomial, i.e. D >= m: the reference does not
            // validate MAX_VAR_DEGREE against the factor count (prover.rs:48-56), and with D < m the interpolant
            // through S(0..D) is not the true polynomial, so there the t = 1 term is computed like the others.
            // (sum of products: D >= the longest term, the degree of the round polynomial)
            unsigned true_degree = m;
            if (sop) {
                true_degree = 0;
                for (int t = 0; t < sop->n_terms; t++) true_degree = sop->len[t] > true_degree ? sop->len[t] : true_degree;
            }
            const bool derive_s1 = degree >= 1 && degree >= true_degree;
            Fe claim_next = Fe{};
            if (derive_s1 && (!sharded || ctx->rank == 0)) {
                const El c = round_eval.at(S.data(), r);
                std::memcpy(claim_next.v, c.v, 32);
            }
            const Fe* claim_ptr = derive_s1 ? &claim_next : nullptr;
            e = timed([&] { next_seq(ctx, sharded); return launch_fold_sums(rf, claim_ptr); });
            cur_len /= 2;
        }
        if (e != cudaSuccess) return cuda_fail(ctx, e, "fold_round_poly");
        st = finish_reduction(ctx, field, np, S.data(), sharded);
        if (st != ZK_OK) return st;
    }
    // last fold (prover.rs:64 in the final iteration): 2 -> 1 entries per factor
    if (n > 0) {
        if (sharded) {  // only reachable when world > 1 and the table never got small enough: gather now
            st = gather();
            if (st != ZK_OK) return st;
        }
        cudaError_t e = zk::launch_fold(field, cur, (int)m, cur_len / 2, fe_from_u64x4(r.v), ctx->stream, &ctx->launches);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "final fold");
        cur_len /= 2;
    }
    if (final_evals_out) {
        for (unsigned k = 0; k < m; k++)
            CU(ctx, cudaMemcpyAsync(final_evals_out + 4 * (size_t)k, cur.t[k], 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i + 1 < ev; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->events[i], ctx->events[i + 1]) == cudaSuccess) {
            ctx->round_ms.push_back(ms);
            ctx->prove_ms[2] += ms;
        }
    }
    count(ctx);
    ctx->prove_ms[0] = timer.ms();
    return ZK_OK;
}

int zk_sumcheck_prove(zk_ctx* ctx, zk_table* const* tables, unsigned m, unsigned degree, const uint64_t sum[4],
                      int absorb_initial_poly, uint64_t* round_polys_out, uint64_t* challenges_out,
                      uint64_t* final_evals_out) {
    return prove_core(ctx, tables, m, degree, sum, absorb_initial_poly, round_polys_out, challenges_out, final_evals_out,
                      nullptr);
}

int zk_sumcheck_prove_host(zk_ctx* ctx, int field, const uint64_t* const* host_tables, unsigned m, unsigned n_vars,
                           unsigned degree, const uint64_t* sum, int absorb_initial_poly, uint64_t* round_polys_out,
                           uint64_t* challenges_out, uint64_t* final_evals_out, uint64_t sum_out[4]) {
    if (!ctx || !valid_field(field)) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (m == 0 || !host_tables) return fail(ctx, ZK_ERR_EMPTY_PRODUCT);
    if (m > ZK_MAX_FACTORS) return fail(ctx, ZK_ERR_UNSUPPORTED, "more than ZK_MAX_FACTORS factors");
    if (n_vars >= 40) return fail(ctx, ZK_ERR_INVALID_ARG);
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<zk_table*> tabs(m, nullptr);
    auto free_all = [&]() { for (auto t : tabs) delete t; };  // the device memory stays in ctx->host_prove_buf
    const uint64_t len = (uint64_t)1 << n_vars, world = (uint64_t)ctx->world;
    if (world > 1 && len < world) return fail(ctx, ZK_ERR_UNSUPPORTED, "table smaller than the number of ranks");
    int st = ZK_OK;
    // All uploads are queued before any compute; no intermediate synchronisation.  Every table is split into S slices
    // copied on S streams (several DMA engines in flight; measured 147-155 ms against 186-204 ms per 6.4 GB proof on
    // a box whose single-stream rate was 33-36 GB/s); the library stream then waits for all of them.
    // ZK_B200_H2D_STREAMS=S overrides the default of 2.  Sharded: the caller passes this rank's shard (entries rank, rank+world, ... stored densely).
    static const int n_copy = [] {
        const char* e = std::getenv("ZK_B200_H2D_STREAMS");
        int v = e ? std::atoi(e) : 2;
        return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    while (n_copy > 1 && (int)ctx->copy_streams.size() < n_copy) {
        cudaStream_t s = nullptr;
        CU(ctx, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        ctx->copy_streams.push_back(s);
    }
    if (n_copy > 1 && !ctx->copy_done) CU(ctx, cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    const uint64_t local_len = len / world;
    for (unsigned k = 0; k < m && st == ZK_OK; k++) {
        if (!host_tables[k]) { st = fail(ctx, ZK_ERR_INVALID_ARG, "null table"); break; }
        if (ctx->host_prove_buf.size() <= k) { ctx->host_prove_buf.push_back(nullptr); ctx->host_prove_cap.push_back(0); }
        if (ctx->host_prove_cap[k] < local_len) {
            cudaFree(ctx->host_prove_buf[k]);
            ctx->host_prove_buf[k] = nullptr;
            ctx->host_prove_cap[k] = 0;
            cudaError_t ea = cudaMalloc((void**)&ctx->host_prove_buf[k], (size_t)local_len * sizeof(Fe));
            if (ea != cudaSuccess) { cudaGetLastError(); st = cuda_fail(ctx, ea, "cudaMalloc(table)"); break; }
            ctx->host_prove_cap[k] = (size_t)local_len;
        }
        tabs[k] = new (std::nothrow) zk_table{ctx, field, n_vars, local_len, ctx->host_prove_buf[k], ctx->host_prove_cap[k]};
        if (!tabs[k]) { st = fail(ctx, ZK_ERR_OOM); break; }
        cudaError_t e = cudaSuccess;
        if (n_copy == 1 || local_len < (uint64_t)n_copy * 4096) {
            e = cudaMemcpyAsync(tabs[k]->data, host_tables[k], (size_t)local_len * 32, cudaMemcpyHostToDevice, ctx->stream);
        } else {
            const uint64_t slice = local_len / n_copy;
            for (int s = 0; s < n_copy && e == cudaSuccess; s++) {
                const uint64_t lo = s * slice, cnt = (s + 1 == n_copy) ? local_len - lo : slice;
                e = cudaMemcpyAsync(tabs[k]->data + lo, host_tables[k] + lo * 4, (size_t)cnt * 32, cudaMemcpyHostToDevice, ctx->copy_streams[s]);
            }
        }
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "upload");
    }
    if (st == ZK_OK && n_copy > 1) {
        for (cudaStream_t s : ctx->copy_streams) {
            cudaError_t e = cudaEventRecord(ctx->copy_done, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0);
            if (e != cudaSuccess) { st = cuda_fail(ctx, e, "upload join"); break; }
        }
    }
    uint64_t claim[4];
    if (st == ZK_OK) {
        if (sum) std::memcpy(claim, sum, 32);
        else st = zk_product_sum(ctx, tabs.data(), m, claim);
    }
    if (st == ZK_OK && sum_out) std::memcpy(sum_out, claim, 32);
    if (st == ZK_OK)
        st = zk_sumcheck_prove(ctx, tabs.data(), m, degree, claim, absorb_initial_poly, round_polys_out, challenges_out,
                               final_evals_out);
    free_all();
    return st;
}

// ---- sum of products (beyond the reference's ProductPoly) --------------------------------
namespace {

// Validates (tables, terms) and builds the kernels' SopSpec.  term_len[t] factors of term t follow each other in
// term_factors; every factor is an index into tables[].
int sop_spec_from(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                  const uint8_t* term_factors, unsigned n_terms, zk::SopSpec* spec) {
    int st = product_check(ctx, tables, n_tables, true);
    if (st != ZK_OK) return st;
    for (unsigned a = 0; a < n_tables; a++)
        for (unsigned b = a + 1; b < n_tables; b++)
            if (tables[a] == tables[b] || tables[a]->data == tables[b]->data)
                return fail(ctx, ZK_ERR_INVALID_ARG, "a table is listed twice: list it once and repeat its index in the terms");
    if (n_terms == 0 || !term_len || !term_factors) return fail(ctx, ZK_ERR_EMPTY_PRODUCT);
    if (n_terms > (unsigned)zk::kMaxTerms) return fail(ctx, ZK_ERR_UNSUPPORTED, "more than 8 terms");
    *spec = zk::SopSpec{};
    spec->n_tables = (int)n_tables;
    spec->n_terms = (int)n_terms;
    size_t off = 0;
    for (unsigned t = 0; t < n_terms; t++) {
        if (term_len[t] == 0) return fail(ctx, ZK_ERR_EMPTY_PRODUCT);
        if (term_len[t] > ZK_MAX_FACTORS) return fail(ctx, ZK_ERR_UNSUPPORTED, "more than ZK_MAX_FACTORS factors in a term");
        spec->len[t] = term_len[t];
        for (unsigned i = 0; i < term_len[t]; i++) {
            if (term_factors[off + i] >= n_tables) return fail(ctx, ZK_ERR_INVALID_ARG, "term factor index out of range");
            spec->fac[t][i] = term_factors[off + i];
        }
        off += term_len[t];
    }
    return ZK_OK;
}

}  // namespace

int zk_sop_combine(int field, const uint8_t* term_len, const uint8_t* term_factors, unsigned n_terms,
                   const uint64_t* table_values, unsigned n_tables, uint64_t out[4]) {
    if (!valid_field(field) || !term_len || !term_factors || !table_values || !out || n_terms == 0) return ZK_ERR_INVALID_ARG;
    const Field F(field);
    El acc = F.zero();
    size_t off = 0;
    for (unsigned t = 0; t < n_terms; t++) {
        if (term_len[t] == 0) return ZK_ERR_INVALID_ARG;
        El pr = F.one();
        for (unsigned i = 0; i < term_len[t]; i++) {
            if (term_factors[off + i] >= n_tables) return ZK_ERR_INVALID_ARG;
            pr = F.mul(pr, el_from(table_values + 4 * (size_t)term_factors[off + i]));
        }
        acc = F.add(acc, pr);
        off += term_len[t];
    }
    std::memcpy(out, acc.v, 32);
    return ZK_OK;
}

int zk_sop_round_poly(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                      const uint8_t* term_factors, unsigned n_terms, unsigned degree, uint64_t* out) {
    if (!ctx || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    zk::SopSpec spec;
    int st = sop_spec_from(ctx, tables, n_tables, term_len, term_factors, n_terms, &spec);
    if (st != ZK_OK) return st;
    if (!zk::sop_degree_supported((int)degree)) return fail(ctx, ZK_ERR_UNSUPPORTED, "sum-of-products rounds support MAX_VAR_DEGREE 1..4");
    if (tables[0]->n_vars == 0 || tables[0]->local_len < 2) return fail(ctx, ZK_ERR_VAR_RANGE);
    CU(ctx, cudaSetDevice(ctx->device));
    next_seq(ctx, true);
    CU(ctx, zk::launch_sop_round_poly(tables[0]->field, ptrs_of(tables, n_tables), spec, (int)degree, tables[0]->local_len / 2,
                                      ctx->scratch, ctx->stream, &ctx->launches));
    st = finish_reduction(ctx, tables[0]->field, (int)degree + 1, out, true);
    count(ctx);
    return st;
}

int zk_sop_sum(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
               const uint8_t* term_factors, unsigned n_terms, uint64_t out[4]) {
    if (!ctx || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, n_tables, true);
    if (st != ZK_OK) return st;
    if (tables[0]->n_vars == 0) {  // a constant: the sum over the empty hypercube is the value itself
        if (ctx->world > 1) return fail(ctx, ZK_ERR_UNSUPPORTED, "zero-variable tables on a sharded context");
        zk::SopSpec spec;
        st = sop_spec_from(ctx, tables, n_tables, term_len, term_factors, n_terms, &spec);
        if (st != ZK_OK) return st;
        CU(ctx, cudaSetDevice(ctx->device));
        std::vector<uint64_t> vals((size_t)n_tables * 4);
        for (unsigned k = 0; k < n_tables; k++)
            CU(ctx, cudaMemcpyAsync(vals.data() + 4 * (size_t)k, tables[k]->data, 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return zk_sop_combine(tables[0]->field, term_len, term_factors, n_terms, vals.data(), n_tables, out);
    }
    // S(0) + S(1) of the first round polynomial is the sum over the whole hypercube
    uint64_t s01[8];
    st = zk_sop_round_poly(ctx, tables, n_tables, term_len, term_factors, n_terms, 1, s01);
    if (st != ZK_OK) return st;
    const Field F(tables[0]->field);
    const El sum = F.add(el_from(s01), el_from(s01 + 4));
    std::memcpy(out, sum.v, 32);
    return ZK_OK;
}

int zk_sop_evaluate(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                    const uint8_t* term_factors, unsigned n_terms, const uint64_t* point, unsigned len, uint64_t out[4]) {
    if (!ctx || !out) return fail(ctx, ZK_ERR_INVALID_ARG);
    zk::SopSpec spec;
    int st = sop_spec_from(ctx, tables, n_tables, term_len, term_factors, n_terms, &spec);
    if (st != ZK_OK) return st;
    if (len != tables[0]->n_vars) return fail(ctx, ZK_ERR_EVALUATE_ARITY);
    std::vector<uint64_t> vals((size_t)n_tables * 4);
    for (unsigned k = 0; k < n_tables; k++) {
        st = zk_mle_evaluate(ctx, tables[k], point, len, vals.data() + 4 * (size_t)k);
        if (st != ZK_OK) return st;
    }
    return zk_sop_combine(tables[0]->field, term_len, term_factors, n_terms, vals.data(), n_tables, out);
}

int zk_sumcheck_prove_sop(zk_ctx* ctx, zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                          const uint8_t* term_factors, unsigned n_terms, unsigned degree, const uint64_t sum[4],
                          int absorb_initial_poly, uint64_t* round_polys_out, uint64_t* challenges_out,
                          uint64_t* final_evals_out) {
    if (!ctx || !sum) return fail(ctx, ZK_ERR_INVALID_ARG);
    zk::SopSpec spec;
    int st = sop_spec_from(ctx, tables, n_tables, term_len, term_factors, n_terms, &spec);
    if (st != ZK_OK) return st;
    if (!zk::sop_degree_supported((int)degree)) return fail(ctx, ZK_ERR_UNSUPPORTED, "sum-of-products rounds support MAX_VAR_DEGREE 1..4");
    return prove_core(ctx, tables, n_tables, degree, sum, absorb_initial_poly, round_polys_out, challenges_out,
                      final_evals_out, &spec);
}

// ---- verifier -------------------------------------------------------------------------------------
namespace {

// UnivariatePolynomial::interpolate over x = 0..D (univariate_poly.rs:43-80), coefficient form.
void interpolate(const Field& F, const El* ys, int n, El* coef) {
    std::vector<El> xs(n);
    for (int i = 0; i < n; i++) xs[i] = F.from_u64((uint64_t)i);
    for (int i = 0; i < n; i++) coef[i] = F.zero();
    std::vector<El> basis(n + 1), tmp(n + 1);
    for (int li = 0; li < n; li++) {
        int nb = 1;
        basis[0] = F.one();
        for (int xi = 0; xi < n; xi++) {
            if (xi == li) continue;
            El den = F.inverse(F.sub(xs[li], xs[xi]));
            El c0 = F.mul(F.neg(xs[xi]), den), c1 = den;  // (x - x_i) / (x_l - x_i)
            for (int i = 0; i <= nb; i++) tmp[i] = F.zero();
            for (int i = 0; i < nb; i++) {
                tmp[i] = F.add(tmp[i], F.mul(basis[i], c0));
                tmp[i + 1] = F.add(tmp[i + 1], F.mul(basis[i], c1));
            }
            nb++;
            for (int i = 0; i < nb; i++) basis[i] = tmp[i];
        }
        for (int i = 0; i < nb; i++) coef[i] = F.add(coef[i], F.mul(basis[i], ys[li]));
    }
}
El horner(const Field& F, const El* coef, int n, const El& x) {  // univariate_poly.rs:29-40
    El acc = F.zero();
    for (int i = n - 1; i >= 0; i--) acc = F.add(F.mul(acc, x), coef[i]);
    return acc;
}

// verifier.rs:44-78
int verify_internal(const Field& F, zk::host::Transcript& tr, const uint64_t sum[4], const uint64_t* round_polys,
                    unsigned n_rounds, unsigned degree, El* subclaim_sum, uint64_t* challenges_out) {
    const int np = (int)degree + 1;
    El claimed = el_from(sum);
    tr.append_element(F, claimed);  // :50
    std::vector<El> rp(np), coef(np);
    for (unsigned r = 0; r < n_rounds; r++) {
        for (int t = 0; t < np; t++) {
            rp[t] = el_from(round_polys + 4 * ((size_t)r * np + t));
            tr.append_element(F, rp[t]);  // :56
        }
        interpolate(F, rp.data(), np, coef.data());  // :58
        El p0 = horner(F, coef.data(), np, F.zero()), p1 = horner(F, coef.data(), np, F.one());
        if (claimed != F.add(p0, p1)) return ZK_ERR_ROUND_CHECK;  // :64-66
        El ch = tr.sample_field_element(F);                        // :69
        claimed = horner(F, coef.data(), np, ch);                  // :70
        if (challenges_out) std::memcpy(challenges_out + 4 * (size_t)r, ch.v, 32);
    }
    *subclaim_sum = claimed;
    return ZK_OK;
}

}  // namespace

int zk_sumcheck_verify_partial(int field, const uint64_t sum[4], const uint64_t* round_polys, unsigned n_rounds,
                               unsigned degree, uint64_t subclaim_sum_out[4], uint64_t* challenges_out) {
    if (!valid_field(field) || !sum || (n_rounds && !round_polys) || !subclaim_sum_out || degree > ZK_MAX_DEGREE)
        return ZK_ERR_INVALID_ARG;
    Field F(field);
    zk::host::Transcript tr;
    El sub;
    int st = verify_internal(F, tr, sum, round_polys, n_rounds, degree, &sub, challenges_out);
    if (st != ZK_OK) return st;
    std::memcpy(subclaim_sum_out, sub.v, 32);
    return ZK_OK;
}

int zk_sumcheck_verify(zk_ctx* ctx, const zk_table* const* tables, unsigned m, const uint64_t sum[4],
                       const uint64_t* round_polys, unsigned n_rounds, unsigned degree) {
    if (!ctx || !sum || (n_rounds && !round_polys)) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, false);
    if (st != ZK_OK) return st;
    if (degree > ZK_MAX_DEGREE) return fail(ctx, ZK_ERR_UNSUPPORTED, "degree > ZK_MAX_DEGREE");
    if (n_rounds != tables[0]->n_vars) return fail(ctx, ZK_ERR_PROOF_ROUNDS);  // verifier.rs:17-19
    CU(ctx, cudaSetDevice(ctx->device));
    Field F(tables[0]->field);
    zk::host::Transcript tr;
    st = absorb_tables(ctx, tables, m, tr);  // :21-22
    count(ctx);
    if (st != ZK_OK) return st;
    El sub;
    std::vector<uint64_t> challenges((size_t)n_rounds * 4 + 4);
    st = verify_internal(F, tr, sum, round_polys, n_rounds, degree, &sub, challenges.data());
    if (st != ZK_OK) return fail(ctx, st);
    uint64_t ev[4];
    st = zk_product_evaluate(ctx, tables, m, challenges.data(), n_rounds, ev);  // :28-30
    if (st != ZK_OK) return fail(ctx, ZK_ERR_INITIAL_EVAL);
    if (el_from(ev) != sub) return fail(ctx, ZK_VERIFY_FALSE);  // :32
    return ZK_OK;
}

int zk_sumcheck_verify_sop(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                           const uint8_t* term_factors, unsigned n_terms, const uint64_t sum[4], const uint64_t* round_polys,
                           unsigned n_rounds, unsigned degree) {
    if (!ctx || !sum || (n_rounds && !round_polys)) return fail(ctx, ZK_ERR_INVALID_ARG);
    zk::SopSpec spec;
    int st = sop_spec_from(ctx, tables, n_tables, term_len, term_factors, n_terms, &spec);
    if (st != ZK_OK) return st;
    if (degree > ZK_MAX_DEGREE) return fail(ctx, ZK_ERR_UNSUPPORTED, "degree > ZK_MAX_DEGREE");
    if (n_rounds != tables[0]->n_vars) return fail(ctx, ZK_ERR_PROOF_ROUNDS);  // verifier.rs:17-19
    CU(ctx, cudaSetDevice(ctx->device));
    Field F(tables[0]->field);
    zk::host::Transcript tr;
    st = absorb_tables(ctx, tables, n_tables, tr);  // :21-22, the tables' to_bytes() in tables[] order
    count(ctx);
    if (st != ZK_OK) return st;
    El sub;
    std::vector<uint64_t> challenges((size_t)n_rounds * 4 + 4);
    st = verify_internal(F, tr, sum, round_polys, n_rounds, degree, &sub, challenges.data());
    if (st != ZK_OK) return fail(ctx, st);
    uint64_t ev[4];
    st = zk_sop_evaluate(ctx, tables, n_tables, term_len, term_factors, n_terms, challenges.data(), n_rounds, ev);  // :28-30
    if (st != ZK_OK) return fail(ctx, ZK_ERR_INITIAL_EVAL);
    if (el_from(ev) != sub) return fail(ctx, ZK_VERIFY_FALSE);  // :32
    return ZK_OK;
}

int zk_sumcheck_proof_dump(int field, const uint64_t sum[4], const uint64_t* round_polys, unsigned n_rounds,
                           unsigned degree, const uint64_t* challenges, const uint64_t* final_evals, unsigned m,
                           uint8_t* out, size_t out_cap, size_t* out_len, uint8_t digest_out[32]) {
    if (!valid_field(field) || !sum || (n_rounds && !round_polys) || !out_len) return ZK_ERR_INVALID_ARG;
    const size_t n_elems = 1 + (size_t)n_rounds * (degree + 1) + (challenges ? n_rounds : 0) + (final_evals ? m : 0);
    *out_len = n_elems * 32;
    if (!out) return ZK_OK;
    if (out_cap < *out_len) return ZK_ERR_INVALID_ARG;
    Field F(field);
    uint8_t* w = out;
    auto put = [&](const uint64_t* e) { F.to_be32(el_from(e), w); w += 32; };
    put(sum);
    for (size_t i = 0; i < (size_t)n_rounds * (degree + 1); i++) put(round_polys + 4 * i);
    if (challenges) for (unsigned i = 0; i < n_rounds; i++) put(challenges + 4 * (size_t)i);
    if (final_evals) for (unsigned k = 0; k < m; k++) put(final_evals + 4 * (size_t)k);
    if (digest_out) zk_keccak256(out, *out_len, digest_out);
    return ZK_OK;
}

// ---- transcript -------------------------------------------------------------------------------------
zk_transcript* zk_transcript_new(void) { return new (std::nothrow) zk_transcript(); }
void zk_transcript_free(zk_transcript* t) { delete t; }
void zk_transcript_append(zk_transcript* t, const uint8_t* data, size_t len) {
    if (t && (data || len == 0)) t->t.append(data ? data : (const uint8_t*)"", len);
}
int zk_transcript_sample_field_element(zk_transcript* t, int field, uint64_t out[4]) {
    if (!t || !out || !valid_field(field)) return ZK_ERR_INVALID_ARG;
    Field F(field);
    El e = t->t.sample_field_element(F);
    std::memcpy(out, e.v, 32);
    return ZK_OK;
}
int zk_transcript_sample_n_field_elements(zk_transcript* t, int field, unsigned n, uint64_t* out) {
    for (unsigned i = 0; i < n; i++) {
        int st = zk_transcript_sample_field_element(t, field, out + 4 * (size_t)i);
        if (st != ZK_OK) return st;
    }
    return ZK_OK;
}
void zk_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    zk::host::Keccak256 h;
    h.update(data ? data : (const uint8_t*)"", len);
    h.finalize_reset(out);
}

// ---- NTT ------------------------------------------------------------------------------------------
int zk_ntt(zk_ctx* ctx, zk_table* inout, int inverse) {
    if (!ctx || !inout) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (ctx->world > 1) return fail(ctx, ZK_ERR_UNSUPPORTED, "NTT on a sharded context (replicas only)");
    Field F(inout->field);
    if (inout->n_vars > F.two_adicity()) return fail(ctx, ZK_ERR_NO_ROOT);
    CU(ctx, cudaSetDevice(ctx->device));
    if (inout->n_vars == 0) return ZK_OK;  // fft_internal: len == 1 -> unchanged (ifft scales by 1^-1 = 1)
    zk::NttPlan* plan = nullptr;
    for (auto* pl : ctx->ntt_plans)
        if (zk::ntt_plan_is(pl, inout->field, inout->n_vars, inverse != 0)) plan = pl;
    cudaError_t e = cudaSuccess;
    if (!plan) {
        if (ctx->ntt_plans.size() >= 2) {  // keep at most a forward/inverse pair resident
            for (auto* pl : ctx->ntt_plans) zk::ntt_plan_destroy(pl);
            ctx->ntt_plans.clear();
        }
        e = zk::ntt_plan_create(inout->field, inout->n_vars, inverse != 0, ctx->stream, &plan, &ctx->launches);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "ntt_plan_create");
        ctx->ntt_plans.push_back(plan);
    }
    Fe* result = nullptr;
    e = zk::ntt_execute(plan, inout->data, &result, ctx->stream, &ctx->launches);
    if (e == cudaSuccess && result != inout->data) {
        if (inout->capacity == inout->local_len) {  // swap buffers with the plan: no copy
            zk::ntt_plan_adopt_scratch(plan, inout->data);
            inout->data = result;
        } else {
            e = cudaMemcpyAsync(inout->data, result, (size_t)inout->local_len * 32, cudaMemcpyDeviceToDevice, ctx->stream);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    count(ctx);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "ntt_execute");
    return ZK_OK;
}

int zk_ntt_host(zk_ctx* ctx, int field, uint64_t* data, uint64_t len, int inverse) {
    if (!ctx || !data || !valid_field(field)) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (len == 0 || (len & (len - 1))) return fail(ctx, ZK_ERR_NOT_POW2);
    zk_table* t = nullptr;
    int st = zk_table_upload(ctx, field, data, len, log2_exact(len), &t);
    if (st != ZK_OK) return st;
    st = zk_ntt(ctx, t, inverse);
    if (st == ZK_OK) st = zk_table_download(ctx, t, data);
    zk_table_free(t);
    return st;
}

// ---- multi-GPU NTT ---------------------------------------------------------------------
// G = 2^g ranks, N = 2^n points, M = N / G, C = M / G.  Rank q holds the strided shard a_q[j] = a[j G + q] (this
// library's table sharding); the forward transform leaves the contiguous block X[c M .. (c+1) M) on rank c, the
// inverse maps blocks back to strided shards.  Forward: local M-point NTT (kernels_ntt.cu), twiddle by w_N^(q k'),
// all-to-all of C-entry chunks, G-point DFT across the ranks' values, all-to-all (ntt_sharded_kernels.cuh has the
// algebra; tests/test_ntt_sharded_model.py replays exactly these steps over integers).  The same step functions run
// either on this process's one rank with NCCL send/recv groups as the transport (zk_ntt_sharded) or on G virtual
// ranks of ONE GPU with device-to-device copies as the transport (zk_ntt_virtual_sharded: every kernel and index map
// of the multi-GPU path, testable on a single GPU).
namespace {

struct NttRank {
    int rank;
    Fe** data;       // the owner's buffer pointer (M elements): the local transform may swap it with the plan's scratch
    bool swappable;  // *data was allocated with exactly M elements by cudaMalloc, so it may be swapped
    Fe* B;           // [G][C] landing zone of the first exchange
    Fe* Y;           // [G][C] output of the G-point DFT
};

// the single-GPU transform of one M-element buffer on ctx->stream (plan cache as in zk_ntt), no synchronisation
int sharded_local_ntt(zk_ctx* ctx, int field, unsigned log_m, bool inverse, Fe** data, bool swappable) {
    if (log_m == 0) return ZK_OK;
    zk::NttPlan* plan = nullptr;
    for (auto* pl : ctx->ntt_plans)
        if (zk::ntt_plan_is(pl, field, log_m, inverse)) plan = pl;
    if (!plan) {
        if (ctx->ntt_plans.size() >= 2) {
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            for (auto* pl : ctx->ntt_plans) zk::ntt_plan_destroy(pl);
            ctx->ntt_plans.clear();
        }
        cudaError_t e = zk::ntt_plan_create(field, log_m, inverse, ctx->stream, &plan, &ctx->launches);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "ntt_plan_create");
        ctx->ntt_plans.push_back(plan);
    }
    Fe* result = nullptr;
    cudaError_t e = zk::ntt_execute(plan, *data, &result, ctx->stream, &ctx->launches);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "ntt_execute");
    if (result != *data) {
        if (swappable) {
            zk::ntt_plan_adopt_scratch(plan, *data);
            *data = result;
        } else {
            CU(ctx, cudaMemcpyAsync(*data, result, ((size_t)1 << log_m) * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    return ZK_OK;
}

// All-to-all of `chunk`-element pieces: piece r of rank q's send buffer lands as piece q of rank r's receive buffer.
// One local rank: an NCCL send/recv group over the context's communicator; G local (virtual) ranks: plain copies.
int sharded_exchange(zk_ctx* ctx, int G, uint64_t chunk, const std::vector<NttRank>& ranks, const std::vector<const Fe*>& send,
                     const std::vector<Fe*>& recv) {
    const size_t bytes = (size_t)chunk * sizeof(Fe);
    if (ranks.size() == 1) {
        if (!nccl().p2p_ok || !ctx->comm) return fail(ctx, ZK_ERR_NCCL, "ncclSend/ncclRecv unavailable");
        int rc = nccl().GroupStart();
        for (int r = 0; r < G && rc == 0; r++) {
            rc = nccl().Send(send[0] + (size_t)r * chunk, bytes, kNcclUint8, r, ctx->comm, ctx->stream);
            if (rc == 0) rc = nccl().Recv(recv[0] + (size_t)r * chunk, bytes, kNcclUint8, r, ctx->comm, ctx->stream);
        }
        const int rc_end = nccl().GroupEnd();
        if (rc == 0) rc = rc_end;
        if (rc != 0) return fail(ctx, ZK_ERR_NCCL, nccl().GetErrorString ? nccl().GetErrorString(rc) : "send/recv");
        return ZK_OK;
    }
    for (int q = 0; q < G; q++)
        for (int r = 0; r < G; r++)
            CU(ctx, cudaMemcpyAsync(recv[(size_t)r] + (size_t)q * chunk, send[(size_t)q] + (size_t)r * chunk, bytes,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    return ZK_OK;
}

// x[k'] *= w_N^(+-q k') on rank q (tables: t_lo 2^lo_bits entries, t_hi M >> lo_bits entries, rebuilt per rank)
int sharded_twiddle(zk_ctx* ctx, const Field& F, int field, unsigned n, unsigned log_m, bool inverse, int q, Fe* x, Fe* t_lo,
                    Fe* t_hi, unsigned lo_bits) {
    if (q == 0) return ZK_OK;
    El w = F.root_of_unity(n);
    if (inverse) w = F.inverse(w);
    const uint64_t e[1] = {(uint64_t)q};
    w = F.pow(w, e, 1);
    const Fe wf = fe_from_u64x4(w.v);
    const uint64_t m = (uint64_t)1 << log_m;
    CU(ctx, zk::launch_pow_table(field, t_lo, (uint64_t)1 << lo_bits, wf, 0, ctx->stream, &ctx->launches));
    CU(ctx, zk::launch_pow_table(field, t_hi, m >> lo_bits, wf, lo_bits, ctx->stream, &ctx->launches));
    CU(ctx, zk::launch_twiddle_mul(field, x, m, t_lo, t_hi, lo_bits, ctx->stream, &ctx->launches));
    return ZK_OK;
}

// The whole factorised transform over the given local ranks (one real rank, or all G virtual ones).
int sharded_ntt_run(zk_ctx* ctx, int field, unsigned n, int G, bool inverse, std::vector<NttRank>& ranks, Fe* t_lo, Fe* t_hi,
                    unsigned lo_bits) {
    const Field F(field);
    unsigned g = 0;
    while ((1 << g) < G) g++;
    const unsigned log_m = n - g;
    const uint64_t M = (uint64_t)1 << log_m, C = M >> g;
    // w_G^(+-i), i < G/2, and G^-1
    El wg = F.root_of_unity(g);
    if (inverse) wg = F.inverse(wg);
    Fe w_half[4];
    El p = F.one();
    for (int i = 0; i < G / 2; i++) {
        w_half[i] = fe_from_u64x4(p.v);
        p = F.mul(p, wg);
    }
    const El ginv = F.inverse(F.from_u64((uint64_t)G));
    const Fe scale = fe_from_u64x4(ginv.v);
    std::vector<const Fe*> send(ranks.size());
    std::vector<Fe*> recv(ranks.size());
    int st = ZK_OK;
    if (!inverse) {
        for (auto& r : ranks) {
            st = sharded_local_ntt(ctx, field, log_m, false, r.data, r.swappable);
            if (st == ZK_OK) st = sharded_twiddle(ctx, F, field, n, log_m, false, r.rank, *r.data, t_lo, t_hi, lo_bits);
            if (st != ZK_OK) return st;
        }
    }
    for (size_t i = 0; i < ranks.size(); i++) { send[i] = *ranks[i].data; recv[i] = ranks[i].B; }
    st = sharded_exchange(ctx, G, C, ranks, send, recv);
    if (st != ZK_OK) return st;
    for (auto& r : ranks)
        CU(ctx, zk::launch_gdft(field, G, r.B, r.Y, C, w_half, inverse ? &scale : nullptr, ctx->stream, &ctx->launches));
    for (size_t i = 0; i < ranks.size(); i++) { send[i] = ranks[i].Y; recv[i] = *ranks[i].data; }
    st = sharded_exchange(ctx, G, C, ranks, send, recv);
    if (st != ZK_OK) return st;
    if (inverse) {
        for (auto& r : ranks) {
            st = sharded_twiddle(ctx, F, field, n, log_m, true, r.rank, *r.data, t_lo, t_hi, lo_bits);
            if (st == ZK_OK) st = sharded_local_ntt(ctx, field, log_m, true, r.data, r.swappable);
            if (st != ZK_OK) return st;
        }
    }
    return ZK_OK;
}

inline unsigned sharded_lo_bits(unsigned log_m) { return log_m < 13 ? log_m : 13; }

int sharded_ntt_check(zk_ctx* ctx, const zk_table* t, int G) {
    if (G != 2 && G != 4 && G != 8) return fail(ctx, ZK_ERR_UNSUPPORTED, "multi-GPU NTT: 2, 4 or 8 ranks");
    const Field F(t->field);
    if (t->n_vars > F.two_adicity()) return fail(ctx, ZK_ERR_NO_ROOT);
    unsigned g = 0;
    while ((1 << g) < G) g++;
    if (t->n_vars < 2 * g) return fail(ctx, ZK_ERR_UNSUPPORTED, "multi-GPU NTT needs at least ranks^2 points");
    return ZK_OK;
}

}  // namespace

int zk_ntt_sharded(zk_ctx* ctx, zk_table* inout, int inverse) {
    if (!ctx || !inout) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (ctx->world == 1) return zk_ntt(ctx, inout, inverse);
    const int G = ctx->world;
    int st = sharded_ntt_check(ctx, inout, G);
    if (st != ZK_OK) return st;
    const uint64_t M = ((uint64_t)1 << inout->n_vars) / (uint64_t)G;
    if (inout->local_len != M) return fail(ctx, ZK_ERR_INVALID_ARG, "table is not sharded over this context");
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned log_m = 0;
    while (((uint64_t)1 << log_m) < M) log_m++;
    const unsigned lo_bits = sharded_lo_bits(log_m);
    const size_t n_lo = (size_t)1 << lo_bits, n_hi = (size_t)(M >> lo_bits);
    const size_t need = 2 * (size_t)M + n_lo + n_hi;
    if (ctx->gather_cap < need) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->gather_buf);
        ctx->gather_buf = nullptr;
        ctx->gather_cap = 0;
        CU(ctx, cudaMalloc((void**)&ctx->gather_buf, need * sizeof(Fe)));
        ctx->gather_cap = need;
    }
    Fe* base = ctx->gather_buf;
    std::vector<NttRank> ranks(1);
    ranks[0] = NttRank{ctx->rank, &inout->data, inout->capacity == (size_t)M, base, base + M};
    st = sharded_ntt_run(ctx, inout->field, inout->n_vars, G, inverse != 0, ranks, base + 2 * M, base + 2 * M + n_lo, lo_bits);
    count(ctx);
    if (st != ZK_OK) return st;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_ntt_virtual_sharded(zk_ctx* ctx, zk_table* inout, unsigned n_ranks, int inverse) {
    if (!ctx || !inout) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (ctx->world > 1) return fail(ctx, ZK_ERR_UNSUPPORTED, "virtual ranks run on an unsharded context");
    if (n_ranks == 1) return zk_ntt(ctx, inout, inverse);
    const int G = (int)n_ranks;
    int st = sharded_ntt_check(ctx, inout, G);
    if (st != ZK_OK) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t N = (uint64_t)1 << inout->n_vars, M = N / (uint64_t)G;
    unsigned log_m = 0;
    while (((uint64_t)1 << log_m) < M) log_m++;
    const unsigned lo_bits = sharded_lo_bits(log_m);
    const size_t n_lo = (size_t)1 << lo_bits, n_hi = (size_t)(M >> lo_bits);
    // scratch: per rank B and Y, a staging area of N elements, the twiddle tables
    Fe* scratch = nullptr;
    CU(ctx, cudaMalloc((void**)&scratch, ((size_t)3 * N + n_lo + n_hi) * sizeof(Fe)));
    Fe* staging = scratch + 2 * N;
    Fe* t_lo = staging + N;
    std::vector<Fe*> shard((size_t)G, nullptr);
    auto release = [&]() {
        cudaStreamSynchronize(ctx->stream);
        for (Fe* p : shard) cudaFree(p);
        cudaFree(scratch);
    };
    for (int q = 0; q < G; q++) {
        cudaError_t e = cudaMalloc((void**)&shard[(size_t)q], (size_t)M * sizeof(Fe));
        if (e != cudaSuccess) { release(); cudaGetLastError(); return cuda_fail(ctx, e, "cudaMalloc(virtual shard)"); }
    }
    cudaError_t e = cudaSuccess;
    if (!inverse) {  // strided shards of the input
        e = zk::launch_deinterleave(inout->data, staging, M, (unsigned)G, ctx->stream, &ctx->launches);
    } else {  // contiguous blocks of the input
        e = cudaMemcpyAsync(staging, inout->data, (size_t)N * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream);
    }
    for (int q = 0; q < G && e == cudaSuccess; q++)
        e = cudaMemcpyAsync(shard[(size_t)q], staging + (size_t)q * M, (size_t)M * sizeof(Fe), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "virtual shards"); }
    std::vector<NttRank> ranks((size_t)G);
    for (int q = 0; q < G; q++)
        ranks[(size_t)q] = NttRank{q, &shard[(size_t)q], true, scratch + (size_t)q * 2 * M, scratch + (size_t)q * 2 * M + M};
    st = sharded_ntt_run(ctx, inout->field, inout->n_vars, G, inverse != 0, ranks, t_lo, t_lo + n_lo, lo_bits);
    if (st != ZK_OK) { release(); return st; }
    // forward: rank c holds block c of the output; inverse: rank q holds the strided shard q
    for (int q = 0; q < G && e == cudaSuccess; q++)
        e = cudaMemcpyAsync((inverse ? staging : inout->data) + (size_t)q * M, shard[(size_t)q], (size_t)M * sizeof(Fe),
                            cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess && inverse) e = zk::launch_interleave(staging, inout->data, M, (unsigned)G, ctx->stream, &ctx->launches);
    count(ctx);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    release();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "virtual sharded ntt");
    return ZK_OK;
}

int zk_table_upload_local(zk_ctx* ctx, int field, const uint64_t* local_mont_aos, uint64_t local_len, unsigned n_vars,
                          zk_table** out) {
    if (!ctx || !out || !valid_field(field) || (local_len && !local_mont_aos)) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (n_vars >= 40) return fail(ctx, ZK_ERR_INVALID_ARG);
    if (local_len * (uint64_t)ctx->world != ((uint64_t)1 << n_vars)) return fail(ctx, ZK_ERR_EVAL_LEN);
    CU(ctx, cudaSetDevice(ctx->device));
    int st = table_alloc(ctx, field, n_vars, local_len, out);
    if (st != ZK_OK) return st;
    cudaError_t e = cudaMemcpyAsync((*out)->data, local_mont_aos, (size_t)local_len * 32, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        zk_table_free(*out);
        *out = nullptr;
        return cuda_fail(ctx, e, "upload");
    }
    return ZK_OK;
}

// ---- field helpers ----------------------------------------------------------------------------------
int zk_field_from_canonical(int field, const uint64_t* canon, uint64_t* mont, size_t n) {
    if (!valid_field(field) || (n && (!canon || !mont))) return ZK_ERR_INVALID_ARG;
    Field F(field);
    for (size_t i = 0; i < n; i++) {
        uint64_t c[4];
        std::memcpy(c, canon + 4 * i, 32);
        while (F.geq_p(c)) F.sub_p(c);
        El r = F.from_canonical(c);
        std::memcpy(mont + 4 * i, r.v, 32);
    }
    return ZK_OK;
}
int zk_field_to_canonical(int field, const uint64_t* mont, uint64_t* canon, size_t n) {
    if (!valid_field(field) || (n && (!canon || !mont))) return ZK_ERR_INVALID_ARG;
    Field F(field);
    for (size_t i = 0; i < n; i++) F.to_canonical(el_from(mont + 4 * i), canon + 4 * i);
    return ZK_OK;
}
int zk_field_from_u64(int field, uint64_t x, uint64_t out[4]) {
    if (!valid_field(field) || !out) return ZK_ERR_INVALID_ARG;
    El r = Field(field).from_u64(x);
    std::memcpy(out, r.v, 32);
    return ZK_OK;
}
int zk_field_to_bytes_be(int field, const uint64_t* mont, size_t n, uint8_t* out) {
    if (!valid_field(field) || (n && (!mont || !out))) return ZK_ERR_INVALID_ARG;
    Field F(field);
    for (size_t i = 0; i < n; i++) F.to_be32(el_from(mont + 4 * i), out + 32 * i);
    return ZK_OK;
}
int zk_field_from_be_bytes_mod_order(int field, const uint8_t in[32], uint64_t out[4]) {
    if (!valid_field(field) || !in || !out) return ZK_ERR_INVALID_ARG;
    El r = Field(field).from_be32_mod_order(in);
    std::memcpy(out, r.v, 32);
    return ZK_OK;
}
#define ZK_BINOP(name, op)                                                                   \
    int name(int field, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {         \
        if (!valid_field(field) || !a || !b || !out) return ZK_ERR_INVALID_ARG;              \
        El r = Field(field).op(el_from(a), el_from(b));                                      \
        std::memcpy(out, r.v, 32);                                                           \
        return ZK_OK;                                                                        \
    }
ZK_BINOP(zk_field_mul, mul)
ZK_BINOP(zk_field_add, add)
ZK_BINOP(zk_field_sub, sub)
int zk_field_inverse(int field, const uint64_t a[4], uint64_t out[4]) {
    if (!valid_field(field) || !a || !out) return ZK_ERR_INVALID_ARG;
    Field F(field);
    El x = el_from(a);
    if (x == F.zero()) return ZK_ERR_INVALID_ARG;
    El r = F.inverse(x);
    std::memcpy(out, r.v, 32);
    return ZK_OK;
}
int zk_field_root_of_unity(int field, uint64_t n, uint64_t out[4]) {
    if (!valid_field(field) || !out) return ZK_ERR_INVALID_ARG;
    Field F(field);
    if (n == 0 || (n & (n - 1))) return ZK_ERR_NO_ROOT;
    unsigned l = log2_exact(n);
    if (l > F.two_adicity()) return ZK_ERR_NO_ROOT;
    El r = F.root_of_unity(l);
    std::memcpy(out, r.v, 32);
    return ZK_OK;
}

int zk_round_poly_evaluate(int field, const uint64_t* ys, unsigned n_points, const uint64_t x[4], uint64_t out[4]) {
    if (!valid_field(field) || !ys || !x || !out || n_points == 0 || n_points > ZK_MAX_DEGREE + 1) return ZK_ERR_INVALID_ARG;
    const Field F(field);
    for (unsigned t = 0; t < n_points; t++)
        if (!F.is_canonical(el_from(ys + 4 * (size_t)t))) return ZK_ERR_INVALID_ARG;
    if (!F.is_canonical(el_from(x))) return ZK_ERR_INVALID_ARG;
    const RoundPolyEvaluator ev(F, (int)n_points);
    const El r = ev.at(ys, el_from(x));
    std::memcpy(out, r.v, 32);
    return ZK_OK;
}

int zk_microbench_run(zk_ctx* ctx, int field, zk_microbench* out) {
    if (!ctx || !out || !valid_field(field)) return fail(ctx, ZK_ERR_INVALID_ARG);
    CU(ctx, cudaSetDevice(ctx->device));
    zk::MicrobenchResult r{};
    CU(ctx, zk::run_microbench(field, &r, ctx->stream));
    out->imad_wide_per_s = r.imad_wide_per_s;
    out->imad_lo_per_s = r.imad_lo_per_s;
    out->iadd3_per_s = r.iadd3_per_s;
    out->mixed_per_s = r.mixed_per_s;
    out->fe_mul_per_s = r.fe_mul_per_s;
    out->copy_gbs = r.copy_gbs;
    out->read_gbs = r.read_gbs;
    out->sm_clock_mhz = r.sm_clock_mhz;
    out->dfma_per_s = r.dfma_per_s;
    out->fe_mul_fixed_per_s = r.fe_mul_fixed_per_s;
    return ZK_OK;
}

}  // extern "C"
