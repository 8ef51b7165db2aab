// api_ntt.cu — the C ABI of fft / ifft (fft/src/lib.rs:4-19): the single-GPU transform with its plan cache, and the
// multi-GPU factorisation (real ranks over NCCL send/recv groups, or virtual ranks on one GPU).
#include "api_internal.h"

using namespace zkapi;

extern "C" {

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
    if (ctx->world > 1) return fail(ctx, ZK_ERR_UNSUPPORTED, "NTT on a sharded context (replicas only)");
    CU(ctx, cudaSetDevice(ctx->device));
    // The device landing buffer stays with the context between calls (grow-only, like zk_sumcheck_prove_host's): a
    // cudaMalloc + cudaFree of the whole vector per call costs milliseconds and a device-wide synchronisation each.
    if (ctx->ntt_host_cap != len) {  // exact size: zk_ntt then swaps it with the plan's scratch instead of copying back
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->ntt_host_buf);
        ctx->ntt_host_buf = nullptr;
        ctx->ntt_host_cap = 0;
        cudaError_t ea = cudaMalloc((void**)&ctx->ntt_host_buf, (size_t)len * sizeof(Fe));
        if (ea != cudaSuccess) { cudaGetLastError(); return cuda_fail(ctx, ea, "cudaMalloc(ntt vector)"); }
        ctx->ntt_host_cap = (size_t)len;
    }
    zk_table t{ctx, field, log2_exact(len), len, ctx->ntt_host_buf, (size_t)len};
    CU(ctx, cudaMemcpyAsync(t.data, data, (size_t)len * 32, cudaMemcpyHostToDevice, ctx->stream));
    int st = zk_ntt(ctx, &t, inverse);
    ctx->ntt_host_buf = t.data;  // zk_ntt may have exchanged the buffer with its plan's scratch
    if (st == ZK_OK) st = zk_table_download(ctx, &t, data);
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

}  // extern "C"
