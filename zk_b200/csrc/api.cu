// api.cu — the C ABI (include/zk_b200.h) over the sm_100a kernels, part 1: contexts (stream, scratch, the NCCL
// symbol table of sharded contexts), tables, MLE and ProductPoly steps, the transcript, field helpers.
// api_sumcheck.cu holds the prover / sum of products / verifier, api_ntt.cu the NTT entry points.
// No CPU fallback exists: every data-path entry launches kernels on the context's GPU.
#include "api_internal.h"

using namespace zkapi;

namespace zkapi {
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
}  // namespace zkapi

namespace zkapi {

const char* const kMessages[] = {
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

int fail(zk_ctx* ctx, int status, const std::string& detail) {
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

void mailbox_teardown(zk_ctx* c);

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
// The in-place / consuming entry points fold every listed table once per launch: the same table (or the same device
// buffer) listed twice would be folded twice.  The reference's ProductPoly owns its factors (`vec![f.clone(), f.clone()]`
// are two buffers), so a caller that wants f * f passes a clone — the mirrors do that for it.
int distinct_check(zk_ctx* ctx, const zk_table* const* tables, unsigned m) {
    for (unsigned a = 0; a < m; a++)
        for (unsigned b = a + 1; b < m; b++)
            if (tables[a] == tables[b] || tables[a]->data == tables[b]->data)
                return fail(ctx, ZK_ERR_INVALID_ARG, "the same table is listed twice: in-place folds need distinct tables (pass a clone)");
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
    if (allreduce && ctx->world > 1 && !ctx->mbox_in_flight) {
        // NCCL fallback (no peer mailboxes): the reducing launch wrote one 32-bit limb per u64 lane (ReduceScratch::lanes)
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
        if (*flag == 0xffffffffu) return fail(ctx, ZK_ERR_NCCL, "a peer rank never delivered its round sums (mailbox timeout)");
        if ((spins & 0x3fff) == 0x3fff) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) {
                // the stream has drained: the flag store must be visible after a synchronisation, otherwise the
                // launch never published (a failed or skipped reduction) and result_host holds stale sums
                if (*flag != want) {
                    CU(ctx, cudaStreamSynchronize(ctx->stream));
                    if (*flag != want) return fail(ctx, ZK_ERR_CUDA, "the reducing launch finished without publishing its result");
                }
                break;
            }
            if (q != cudaErrorNotReady) return cuda_fail(ctx, q, "round kernel");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (out) std::memcpy(out, ctx->scratch.result_host, (size_t)count_elems * 32);
    return ZK_OK;
}

// Peer mailboxes of a sharded context (kernels.h: MailboxArgs): allocate mine, exchange CUDA IPC handles over the NCCL
// communicator, map every peer's.  Collective; every rank ends with the same answer (all ranks mapped everything, or
// nobody uses the mailboxes and the NCCL all-reduce stays).  ZK_B200_MAILBOX=0 keeps NCCL (A/B runs).
int mailbox_setup(zk_ctx* c) {
    const char* env = std::getenv("ZK_B200_MAILBOX");
    uint64_t ok = !(env && env[0] == '0') && c->world <= zk::kMaxRanks ? 1 : 0;
    const size_t bytes = 2 * (size_t)zk::kMaxRanks * sizeof(zk::MailboxSlot);
    unsigned char* hsend = nullptr;
    unsigned char* hall = nullptr;
    cudaIpcMemHandle_t mine_h;
    std::memset(&mine_h, 0, sizeof(mine_h));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(c, cudaMalloc((void**)&hsend, 64));
    CU(c, cudaMalloc((void**)&hall, 64 * (size_t)c->world));
    if (cudaMalloc((void**)&c->mbox_mine, bytes) != cudaSuccess) { c->mbox_mine = nullptr; ok = 0; cudaGetLastError(); }
    if (c->mbox_mine) {
        CU(c, cudaMemsetAsync(c->mbox_mine, 0, bytes, c->stream));
        if (ok && cudaIpcGetMemHandle(&mine_h, c->mbox_mine) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    }
    CU(c, cudaMemcpyAsync(hsend, &mine_h, 64, cudaMemcpyHostToDevice, c->stream));
    if (nccl().AllGather(hsend, hall, 64, kNcclUint8, c->comm, c->stream) != 0) return fail(c, ZK_ERR_NCCL, "allgather(ipc handles)");
    std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
    CU(c, cudaMemcpyAsync(all.data(), hall, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int q = 0; q < c->world && ok; q++) {
        if (q == c->rank) { c->mbox_peer[q] = c->mbox_mine; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[(size_t)q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
        c->mbox_peer[q] = (zk::MailboxSlot*)p;
    }
    // agreement (and the barrier that orders every rank's memset before anyone's first store): sum of the ok flags
    CU(c, cudaMemcpyAsync(c->lanes, &ok, 8, cudaMemcpyHostToDevice, c->stream));
    if (nccl().AllReduce(c->lanes, c->lanes, 1, kNcclUint64, kNcclSum, c->comm, c->stream) != 0) return fail(c, ZK_ERR_NCCL, "allreduce(mailbox agreement)");
    uint64_t sum = 0;
    CU(c, cudaMemcpyAsync(&sum, c->lanes, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(hsend);
    cudaFree(hall);
    c->mbox_ready = (sum == (uint64_t)c->world);
    if (!c->mbox_ready) mailbox_teardown(c);
    return ZK_OK;
}
void mailbox_teardown(zk_ctx* c) {
    for (int q = 0; q < zk::kMaxRanks; q++) {
        if (c->mbox_peer[q] && q != c->rank) cudaIpcCloseMemHandle(c->mbox_peer[q]);
        c->mbox_peer[q] = nullptr;
    }
    if (c->mbox_mine) cudaFree(c->mbox_mine);
    c->mbox_mine = nullptr;
    c->mbox_ready = false;
    cudaGetLastError();
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

}  // namespace zkapi

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
        st = mailbox_setup(c);
        if (st != ZK_OK) {
            std::fprintf(stderr, "zk_ctx_create_sharded: %s\n", c->last_error.c_str());
            zk_ctx_destroy(c);
            *out = nullptr;
            return st;
        }
    }
    return ZK_OK;
}
int zk_ctx_uses_mailbox(const zk_ctx* c) { return c && c->mbox_ready ? 1 : 0; }

void zk_ctx_destroy(zk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    mailbox_teardown(c);
    if (c->comm) nccl().CommDestroy(c->comm);
    for (auto* pl : c->ntt_plans) zk::ntt_plan_destroy(pl);
    cudaFree(c->gather_buf);
    cudaFree(c->eval_buf);
    cudaFree(c->ntt_host_buf);
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
    for (cudaEvent_t ev : c->slice_events) cudaEventDestroy(ev);
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
    next_seq(ctx, true, zk::has_fused_path((int)m, (int)degree));
    CU(ctx, zk::launch_round_poly(tables[0]->field, ptrs_of(tables, m), (int)m, (int)degree, tables[0]->local_len / 2,
                                  ctx->scratch, ctx->stream, &ctx->launches));
    st = finish_reduction(ctx, tables[0]->field, (int)degree + 1, out, true);
    count(ctx);
    return st;
}

int zk_product_fold_inplace(zk_ctx* ctx, zk_table* const* tables, unsigned m, const uint64_t r[4]) {
    if (!ctx || !r) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st == ZK_OK) st = distinct_check(ctx, tables, m);
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
    if (st == ZK_OK) st = distinct_check(ctx, tables, m);
    if (st != ZK_OK) return st;
    if (degree > ZK_MAX_DEGREE) return fail(ctx, ZK_ERR_UNSUPPORTED, "degree > ZK_MAX_DEGREE");
    if (tables[0]->n_vars < 2 || tables[0]->local_len < 4) return fail(ctx, ZK_ERR_VAR_RANGE);
    CU(ctx, cudaSetDevice(ctx->device));
    next_seq(ctx, true, zk::has_fused_path((int)m, (int)degree));
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
