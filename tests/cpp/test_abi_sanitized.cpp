// A walk over the C ABI meant to run under AddressSanitizer + LeakSanitizer + UBSan against the host mock
// (tests/cpp/hostmock): every "device" buffer is host memory there, so an out-of-bounds access by api.cu's
// orchestration or by a replayed kernel source, a double free or a leak of a handle / event / buffer shows up here.
// Results are cross-checked only lightly (round trips, verifier acceptance): parity is the other tests' job.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "zk_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int st__ = (call);                                                            \
        if (st__ != ZK_OK) { std::printf("%s -> %d (%s)\n", #call, st__, zk_status_string(st__)); return 1; } \
    } while (0)

// One rank of a sharded job (ranks are threads; mock_nccl.cpp provides the collectives): sharded ProductPoly and
// sum-of-products provers at two gather thresholds, the multi-GPU NTT forward and back, then everything released.
static int sharded_rank(int rank, int world, const void* nccl_id) {
    zk_ctx* ctx = nullptr;
    CHECK(zk_ctx_create_sharded(0, rank, world, nccl_id, &ctx));
    const unsigned n = 8, d = 3;
    uint64_t sum[4], rp[8 * 4 * 4], ch[8 * 4], fin[4 * 4];
    const uint8_t tl[3] = {2, 2, 3}, tf[7] = {0, 2, 0, 3, 1, 2, 3};
    for (uint64_t thr : {(uint64_t)4096, (uint64_t)2}) {
        CHECK(zk_ctx_set_gather_threshold(ctx, thr));
        zk_table* t[4];
        for (unsigned k = 0; k < 4; k++) CHECK(zk_table_generate(ctx, 0, 11, k, n, &t[k]));
        CHECK(zk_product_sum(ctx, t, 3, sum));
        CHECK(zk_sumcheck_prove(ctx, t, 3, d, sum, 0, rp, ch, fin));
        for (unsigned k = 0; k < 4; k++) zk_table_free(t[k]);
        for (unsigned k = 0; k < 4; k++) CHECK(zk_table_generate(ctx, 0, 11, k, n, &t[k]));
        CHECK(zk_sop_sum(ctx, t, 4, tl, tf, 3, sum));
        CHECK(zk_sumcheck_prove_sop(ctx, t, 4, tl, tf, 3, d, sum, 0, rp, ch, fin));
        uint64_t sub[4], ch2[8 * 4];
        CHECK(zk_sumcheck_verify_partial(0, sum, rp, n, d, sub, ch2));
        for (unsigned k = 0; k < 4; k++) zk_table_free(t[k]);
    }
    for (int field = 0; field < 2; field++) {
        zk_table* a = nullptr;
        CHECK(zk_table_generate(ctx, field, 3, 1, n, &a));
        std::vector<uint64_t> before(4u << n), after(4u << n);
        CHECK(zk_table_download(ctx, a, before.data()));
        CHECK(zk_ntt_sharded(ctx, a, 0));
        CHECK(zk_ntt_sharded(ctx, a, 1));
        CHECK(zk_table_download(ctx, a, after.data()));
        const size_t local = ((size_t)4 << n) / (size_t)world;
        if (std::memcmp(before.data(), after.data(), local * 8) != 0) { std::printf("sharded ntt round trip (rank %d)\n", rank); return 1; }
        zk_table_free(a);
    }
    zk_ctx_destroy(ctx);
    return 0;
}

static int sharded_jobs() {
    for (int world : {2, 4}) {
        unsigned char id[128];
        CHECK(zk_nccl_unique_id(id));
        std::vector<std::thread> th;
        std::vector<int> rc((size_t)world, -1);
        for (int r = 0; r < world; r++) th.emplace_back([&, r] { rc[(size_t)r] = sharded_rank(r, world, id); });
        for (auto& t : th) t.join();
        for (int r = 0; r < world; r++)
            if (rc[(size_t)r] != 0) { std::printf("world %d rank %d failed\n", world, r); return 1; }
    }
    std::printf("SHARDED WALK OK\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::strcmp(argv[1], "sharded") == 0) return sharded_jobs();
    zk_ctx* ctx = nullptr;
    CHECK(zk_ctx_create(0, &ctx));
    for (int field = 0; field < 2; field++) {
        const unsigned n = 7, m = 3, d = 3;
        zk_table* t[4] = {nullptr, nullptr, nullptr, nullptr};
        for (unsigned k = 0; k < 4; k++) CHECK(zk_table_generate(ctx, field, 11, k, n, &t[k]));
        uint64_t sum[4], rp[7 * 4 * 4], ch[7 * 4], fin[4 * 4], ev[4], ev2[4];
        CHECK(zk_product_sum(ctx, t, m, sum));
        CHECK(zk_product_round_poly(ctx, t, m, d, rp));
        // prove (with the initial absorb) on clones, verify on the originals
        zk_table* c[3];
        for (unsigned k = 0; k < m; k++) CHECK(zk_table_clone(ctx, t[k], &c[k]));
        CHECK(zk_sumcheck_prove(ctx, c, m, d, sum, 1, rp, ch, fin));
        for (unsigned k = 0; k < m; k++) zk_table_free(c[k]);
        CHECK(zk_sumcheck_verify(ctx, t, m, sum, rp, n, d));
        // prove_partial, sub-claim against evaluate
        for (unsigned k = 0; k < m; k++) CHECK(zk_table_clone(ctx, t[k], &c[k]));
        CHECK(zk_sumcheck_prove(ctx, c, m, d, sum, 0, rp, ch, fin));
        for (unsigned k = 0; k < m; k++) zk_table_free(c[k]);
        uint64_t sub[4], ch2[7 * 4];
        CHECK(zk_sumcheck_verify_partial(field, sum, rp, n, d, sub, ch2));
        CHECK(zk_product_evaluate(ctx, t, m, ch2, n, ev));
        if (std::memcmp(ev, sub, 32) != 0 || std::memcmp(ch, ch2, sizeof ch) != 0) { std::printf("sub-claim mismatch\n"); return 1; }
        // step API, fold, partial_evaluate, prod_reduce, to_bytes, download
        for (unsigned k = 0; k < m; k++) CHECK(zk_table_clone(ctx, t[k], &c[k]));
        CHECK(zk_product_fold_then_round_poly(ctx, c, m, d, ch, rp));
        CHECK(zk_product_fold_inplace(ctx, c, m, ch + 4));
        for (unsigned k = 0; k < m; k++) zk_table_free(c[k]);
        zk_table *pe = nullptr, *pr = nullptr;
        CHECK(zk_mle_partial_evaluate(ctx, t[0], 2, ch, 3, &pe));
        CHECK(zk_product_prod_reduce(ctx, t, m, &pr));
        std::vector<uint8_t> bytes(32u << n);
        CHECK(zk_mle_to_bytes(ctx, t[0], bytes.data()));
        std::vector<uint64_t> host(4u << n);
        CHECK(zk_table_download(ctx, pr, host.data()));
        CHECK(zk_mle_evaluate(ctx, t[1], ch, n, ev2));
        zk_table_free(pe);
        zk_table_free(pr);
        // sum of products (GKR shape): prove with absorb, verify, evaluate
        const uint8_t tl[3] = {2, 2, 3}, tf[7] = {0, 2, 0, 3, 1, 2, 3};
        CHECK(zk_sop_sum(ctx, t, 4, tl, tf, 3, sum));
        zk_table* c4[4];
        for (unsigned k = 0; k < 4; k++) CHECK(zk_table_clone(ctx, t[k], &c4[k]));
        CHECK(zk_sumcheck_prove_sop(ctx, c4, 4, tl, tf, 3, d, sum, 1, rp, ch, fin));
        for (unsigned k = 0; k < 4; k++) zk_table_free(c4[k]);
        CHECK(zk_sumcheck_verify_sop(ctx, t, 4, tl, tf, 3, sum, rp, n, d));
        CHECK(zk_sop_round_poly(ctx, t, 4, tl, tf, 3, 2, rp));
        // host-table prover twice (the landing buffers stay with the context), different sizes
        for (unsigned nn : {6u, 5u, 6u}) {
            std::vector<std::vector<uint64_t>> ht(m, std::vector<uint64_t>(4u << nn));
            const uint64_t* ptrs[3];
            for (unsigned k = 0; k < m; k++) {
                zk_table* g = nullptr;
                CHECK(zk_table_generate(ctx, field, 5, k, nn, &g));
                CHECK(zk_table_download(ctx, g, ht[k].data()));
                zk_table_free(g);
                ptrs[k] = ht[k].data();
            }
            CHECK(zk_sumcheck_prove_host(ctx, field, ptrs, m, nn, d, nullptr, 0, rp, ch, fin, sum));
        }
        // NTT: in place (small), through the plan's scratch buffer (swap), host buffers, virtual ranks
        for (unsigned nn : {0u, 1u, 3u, 7u, 9u}) {
            zk_table* a = nullptr;
            CHECK(zk_table_generate(ctx, field, 3, 1, nn, &a));
            std::vector<uint64_t> before(4u << nn), after(4u << nn);
            CHECK(zk_table_download(ctx, a, before.data()));
            CHECK(zk_ntt(ctx, a, 0));
            CHECK(zk_ntt(ctx, a, 1));
            CHECK(zk_table_download(ctx, a, after.data()));
            if (before != after) { std::printf("ntt round trip\n"); return 1; }
            CHECK(zk_ntt_host(ctx, field, after.data(), (uint64_t)1 << nn, 0));
            CHECK(zk_ntt_host(ctx, field, after.data(), (uint64_t)1 << nn, 1));
            if (before != after) { std::printf("ntt_host round trip\n"); return 1; }
            for (unsigned G : {2u, 4u, 8u}) {
                unsigned g = G == 2 ? 1 : (G == 4 ? 2 : 3);
                if (nn < 2 * g) continue;
                CHECK(zk_ntt_virtual_sharded(ctx, a, G, 0));
                CHECK(zk_ntt_virtual_sharded(ctx, a, G, 1));
                CHECK(zk_table_download(ctx, a, after.data()));
                if (before != after) { std::printf("virtual sharded ntt round trip G=%u n=%u\n", G, nn); return 1; }
            }
            zk_table_free(a);
        }
        zk_table* loc = nullptr;
        std::vector<uint64_t> raw(4u << 4, 1);
        CHECK(zk_table_upload_local(ctx, field, raw.data(), 16, 4, &loc));
        zk_table_free(loc);
        for (unsigned k = 0; k < 4; k++) zk_table_free(t[k]);
        // refused calls must not leak (or touch anything out of bounds) either
        zk_table* bad = nullptr;
        if (zk_table_upload(ctx, field, raw.data(), 15, 4, &bad) != ZK_ERR_EVAL_LEN) { std::printf("expected ZK_ERR_EVAL_LEN\n"); return 1; }
        {
            zk_table *a = nullptr, *b = nullptr, *out = nullptr;
            CHECK(zk_table_generate(ctx, field, 1, 0, 4, &a));
            CHECK(zk_table_generate(ctx, field, 1, 1, 5, &b));
            zk_table* mixed[2] = {a, b};
            zk_table* twice[2] = {a, a};
            uint64_t o[64], pt[5 * 4] = {0};
            const uint8_t tl1[1] = {2}, tf1[2] = {0, 1}, tf_bad[2] = {0, 7};
            int want[] = {
                zk_product_sum(ctx, mixed, 2, o) == ZK_ERR_NVARS_MISMATCH,
                zk_product_sum(ctx, mixed, 0, o) == ZK_ERR_EMPTY_PRODUCT,
                zk_sumcheck_prove(ctx, mixed, 2, 2, o, 0, o, o, o) == ZK_ERR_NVARS_MISMATCH,
                zk_mle_evaluate(ctx, a, pt, 3, o) == ZK_ERR_EVALUATE_ARITY,
                zk_mle_partial_evaluate(ctx, a, 3, pt, 2, &out) == ZK_ERR_VAR_RANGE,
                zk_product_round_poly(ctx, twice, 1, ZK_MAX_DEGREE + 1, o) == ZK_ERR_UNSUPPORTED,
                zk_sop_round_poly(ctx, twice, 2, tl1, tf1, 1, 2, o) == ZK_ERR_INVALID_ARG,       /* a table listed twice */
                zk_sop_round_poly(ctx, mixed, 1, tl1, tf_bad, 1, 2, o) == ZK_ERR_INVALID_ARG,    /* factor index out of range */
                zk_sop_round_poly(ctx, mixed, 1, tl1, tf1, 1, 9, o) != ZK_OK,
                zk_sumcheck_verify(ctx, mixed, 1, o, o, 3, 2) == ZK_ERR_PROOF_ROUNDS,
                zk_ntt_virtual_sharded(ctx, a, 8, 0) == ZK_ERR_UNSUPPORTED,                        /* 16 points < 8^2 */
                zk_ntt_virtual_sharded(ctx, a, 3, 0) == ZK_ERR_UNSUPPORTED,
                zk_ntt_host(ctx, field, o, 12, 0) == ZK_ERR_NOT_POW2,
                zk_table_upload_local(ctx, field, o, 3, 4, &out) == ZK_ERR_EVAL_LEN,
                zk_sumcheck_prove_host(ctx, field, nullptr, 0, 4, 2, nullptr, 0, o, o, o, o) == ZK_ERR_EMPTY_PRODUCT,
            };
            for (size_t i = 0; i < sizeof(want) / sizeof(want[0]); i++)
                if (!want[i]) { std::printf("refused call %zu returned an unexpected status\n", i); return 1; }
            if (out != nullptr) { std::printf("a refused call produced a table\n"); return 1; }
            zk_table_free(a);
            zk_table_free(b);
        }
    }
    zk_transcript* tr = zk_transcript_new();
    zk_transcript_append(tr, (const uint8_t*)"abc", 3);
    uint64_t r[4];
    CHECK(zk_transcript_sample_field_element(tr, 0, r));
    zk_transcript_free(tr);
    zk_ctx_destroy(ctx);
    std::printf("ABI WALK OK\n");
    return 0;
}
