// api_sumcheck.cu — the C ABI of the sumcheck protocol layer: the Fiat-Shamir round loop of the prover (host Keccak
// between kernel launches; ProductPoly and sum of products share it), the initial-poly absorb pipeline, the verifier,
// the proof dump.  See include/zk_b200.h for the contract of every entry point.
#include "api_internal.h"

using namespace zkapi;

extern "C" {

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
// `round0` (optional, unsharded ProductPoly only): the sums of round 0, already computed by the caller while the tables
// were still arriving (zk_sumcheck_prove_host) — the loop then starts at the first challenge.
struct Round0 {
    const uint64_t* sums;  // (degree + 1) elements
    float kernel_ms;       // device time of the launches that produced them (reported as round 0's)
};
static int prove_core(zk_ctx* ctx, zk_table* const* tables, unsigned m, unsigned degree, const uint64_t sum[4],
                      int absorb_initial_poly, uint64_t* round_polys_out, uint64_t* challenges_out,
                      uint64_t* final_evals_out, const zk::SopSpec* sop, const Round0* round0 = nullptr) {
    if (!ctx || !sum) return fail(ctx, ZK_ERR_INVALID_ARG);
    int st = product_check(ctx, tables, m, true);
    if (st == ZK_OK) st = distinct_check(ctx, tables, m);  // the prover consumes (folds in place) every listed table
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
        // one NCCL group for the m all-gathers (a single fused launch) and one interleave launch for all tables:
        // landing zones [k][q][j] and interleaved tables [k][j G + q] are stored back to back
        Fe* const stage0 = ctx->gather_buf;
        Fe* const full0 = ctx->gather_buf + (size_t)m * L * G;
        const bool grouped = nccl().p2p_ok;
        int rc = grouped ? nccl().GroupStart() : 0;
        for (unsigned k = 0; k < m && rc == 0; k++)
            rc = nccl().AllGather(cur.t[k], stage0 + (size_t)k * L * G, (size_t)L * 32, kNcclUint8, ctx->comm, ctx->stream);
        if (grouped) {
            const int rc_end = nccl().GroupEnd();
            if (rc == 0) rc = rc_end;
        }
        if (rc != 0) return fail(ctx, ZK_ERR_NCCL, "allgather");
        cudaError_t e = zk::launch_interleave(stage0, full0, L, (unsigned)G, ctx->stream, &ctx->launches, m);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "gather");
        for (unsigned k = 0; k < m; k++) cur.t[k] = full0 + (size_t)k * L * G;
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

    // one reducing launch per round (every fused path; the sum of products has no other): it may all-reduce in-kernel
    const bool single_launch = sop != nullptr || zk::has_fused_path((int)m, (int)degree);
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

    if (n > 0 && round0 && !sharded && !sop) {
        std::memcpy(S.data(), round0->sums, (size_t)np * 32);
    } else if (n > 0) {
        if (sharded && (cur_len < 2 || cur_len <= ctx->gather_threshold)) {
            st = gather();
            if (st != ZK_OK) return st;
        }
        cudaError_t e = timed([&] { next_seq(ctx, sharded, single_launch); return launch_sums(); });
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
            e = timed([&] { next_seq(ctx, sharded, single_launch); return launch_sums(); });
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
            e = timed([&] { next_seq(ctx, sharded, single_launch); return launch_fold_sums(rf, claim_ptr); });
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
    if (n > 0 && round0 && ctx->world == 1 && !sop) {
        ctx->round_ms.push_back(round0->kernel_ms);
        ctx->prove_ms[2] += round0->kernel_ms;
    }
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
    }
    // Round 0 overlapped with the upload (one GPU, prove_partial, fused path): the round sums are additive over index
    // ranges, so the tables travel in slices of pairs — the low entries of slice c of every table on one copy stream, the
    // high entries (j + N/2) on the other — and the round-0 kernel runs on slice c as soon as both halves have landed,
    // while slice c + 1 is on the wire.  Only the last slice's kernel is left after the last byte: the 3 ms of round 0
    // disappear behind the 116 ms of PCIe.  The per-slice sums are added on the host (exact modular additions).
    // ZK_B200_H2D_OVERLAP=0 keeps the plain upload-then-prove order.
    static const bool overlap_enabled = [] {
        const char* e = std::getenv("ZK_B200_H2D_OVERLAP");
        return !(e && e[0] == '0');
    }();
    static const unsigned overlap_min_log = [] {  // smallest local table (log2 entries) that is sliced; tests lower it
        const char* e = std::getenv("ZK_B200_H2D_OVERLAP_MIN_LOG");
        const int v = e ? std::atoi(e) : 21;
        return (unsigned)(v < 5 ? 5 : (v > 40 ? 40 : v));
    }();
    const unsigned np = degree + 1;
    const bool overlap = st == ZK_OK && overlap_enabled && world == 1 && !absorb_initial_poly && n_copy >= 2 && n_vars >= 1 &&
                         local_len >= ((uint64_t)1 << overlap_min_log) && degree <= ZK_MAX_DEGREE && zk::has_fused_path((int)m, (int)degree);
    std::vector<uint64_t> S0((size_t)np * 4, 0);
    float round0_ms = 0;
    if (overlap) {
        const Field F(field);
        const uint64_t half = local_len / 2;
        const unsigned n_slices = 16;
        const uint64_t slice = half / n_slices;
        while (ctx->slice_events.size() < 4 * (size_t)n_slices) {
            cudaEvent_t ev = nullptr;
            CU(ctx, cudaEventCreateWithFlags(&ev, ctx->slice_events.size() < 2 * (size_t)n_slices ? cudaEventDisableTiming : cudaEventDefault));
            ctx->slice_events.push_back(ev);
        }
        cudaError_t e = cudaSuccess;
        for (unsigned c = 0; c < n_slices && e == cudaSuccess; c++) {  // every copy is queued up front
            const uint64_t lo = c * slice;
            for (int h = 0; h < 2 && e == cudaSuccess; h++) {
                cudaStream_t cs = ctx->copy_streams[(size_t)h];
                for (unsigned k = 0; k < m && e == cudaSuccess; k++)
                    e = cudaMemcpyAsync(tabs[k]->data + h * half + lo, host_tables[k] + (h * half + lo) * 4, (size_t)slice * 32, cudaMemcpyHostToDevice, cs);
                if (e == cudaSuccess) e = cudaEventRecord(ctx->slice_events[2 * c + h], cs);
            }
        }
        if (e != cudaSuccess) st = cuda_fail(ctx, e, "upload");
        El acc[ZK_MAX_DEGREE + 1];
        for (unsigned t = 0; t < np; t++) acc[t] = F.zero();
        std::vector<uint64_t> part((size_t)np * 4);
        for (unsigned c = 0; c < n_slices && st == ZK_OK; c++) {
            zk::TablePtrs p{};
            for (unsigned k = 0; k < m; k++) p.t[k] = tabs[k]->data + c * slice;
            e = cudaStreamWaitEvent(ctx->stream, ctx->slice_events[2 * c], 0);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->slice_events[2 * c + 1], 0);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->slice_events[2 * n_slices + 2 * c], ctx->stream);
            next_seq(ctx, false);
            if (e == cudaSuccess) e = zk::launch_round_poly_range(field, p, (int)m, (int)degree, slice, half, ctx->scratch, ctx->stream, &ctx->launches);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->slice_events[2 * n_slices + 2 * c + 1], ctx->stream);
            if (e != cudaSuccess) { st = cuda_fail(ctx, e, "round 0 slice"); break; }
            st = finish_reduction(ctx, field, (int)np, part.data(), false);
            for (unsigned t = 0; t < np && st == ZK_OK; t++) acc[t] = F.add(acc[t], el_from(part.data() + 4 * t));
        }
        for (unsigned t = 0; t < np; t++) std::memcpy(S0.data() + 4 * t, acc[t].v, 32);
        for (unsigned c = 0; c < n_slices && st == ZK_OK; c++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->slice_events[2 * n_slices + 2 * c], ctx->slice_events[2 * n_slices + 2 * c + 1]) == cudaSuccess) round0_ms += ms;
        }
    }
    for (unsigned k = 0; k < m && st == ZK_OK && !overlap; k++) {
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
    if (st == ZK_OK && n_copy > 1 && !overlap) {
        for (cudaStream_t s : ctx->copy_streams) {
            cudaError_t e = cudaEventRecord(ctx->copy_done, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0);
            if (e != cudaSuccess) { st = cuda_fail(ctx, e, "upload join"); break; }
        }
    }
    if (overlap) {
        // the claimed sum: the caller's, or the true sum S_0(0) + S_0(1)
        uint64_t claim[4];
        if (st == ZK_OK) {
            if (sum) std::memcpy(claim, sum, 32);
            else {
                const Field F(field);
                const El c = F.add(el_from(S0.data()), el_from(S0.data() + 4));
                std::memcpy(claim, c.v, 32);
            }
            if (sum_out) std::memcpy(sum_out, claim, 32);
            const Round0 r0{S0.data(), round0_ms};
            st = prove_core(ctx, tabs.data(), m, degree, claim, 0, round_polys_out, challenges_out, final_evals_out, nullptr, &r0);
        }
        free_all();
        return st;
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
    // The proof is untrusted input: the reference's proof holds `F` values, which are canonical by construction.  A limb
    // pattern >= p would be hashed as its reduced value but compared as raw limbs (a second encoding of the same
    // element): rejected here instead.
    if (!F.is_canonical(claimed)) return ZK_ERR_INVALID_ARG;
    for (size_t i = 0; i < (size_t)n_rounds * np; i++)
        if (!F.is_canonical(el_from(round_polys + 4 * i))) return ZK_ERR_INVALID_ARG;
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

}  // extern "C"
