/*
 * zk_b200.h — C ABI of the B200-native sumcheck / MLE-fold / NTT path of iammadab/zk.
 *
 * The reference has no FFI: its boundary is the public Rust API of the `polynomial`, `sumcheck`,
 * `transcript` and `fft` crates.  Each entry point below names the reference item it replaces
 * (file:line relative to the reference tree) and keeps its argument meaning and error behaviour;
 * INTEGRATION.md shows the Rust-side `extern "C"` binding a maintainer would add.
 *
 * Conventions
 *  - Field element = 32 bytes = uint64_t[4]: little-endian limbs, Montgomery form (x * 2^256 mod p),
 *    fully reduced.  This is the in-memory layout of ark-ff 0.5 `Fp<MontBackend<_,4>,4>`, so a Rust
 *    `&[Fr]` can be passed as `const uint64_t*` without conversion.
 *  - Every function returns a zk_status.  Where the reference returns `Err(&'static str)` or panics,
 *    the status maps 1:1 to that message; `zk_status_string` / `zk_last_error` return the literal text.
 *  - Nothing aborts or throws across this boundary.  There is no CPU fallback: without a CUDA device
 *    `zk_ctx_create` fails with ZK_ERR_CUDA.
 *  - One host thread per zk_ctx (the reference is single-threaded).  Calls are synchronous.
 */
#ifndef ZK_B200_H
#define ZK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ZK_API __attribute__((visibility("default")))
#else
#define ZK_API
#endif

#define ZK_MAX_FACTORS 8 /* ProductPoly factors supported on the device */
#define ZK_MAX_DEGREE 15 /* MAX_VAR_DEGREE supported */

typedef struct zk_ctx zk_ctx;               /* one GPU (+ its rank in a sharded job) */
typedef struct zk_table zk_table;           /* device-resident MultiLinearPolynomial<F> (evaluation_form.rs:7-10) */
typedef struct zk_transcript zk_transcript; /* transcript::Transcript (transcript/src/lib.rs:5-7) */

typedef enum zk_field {
    ZK_BLS12_381_FR = 0, /* ark_bls12_381::Fr — sumcheck/polynomial tests (sumcheck/src/lib.rs:35) */
    ZK_BLS12_377_FR = 1  /* ark_bls12_377::Fr — the fft test field (fft/src/lib.rs:75) */
} zk_field;

typedef enum zk_status {
    ZK_OK = 0,
    ZK_ERR_EVAL_LEN = 1,       /* "evaluation vec len should equal 2^n_vars"           evaluation_form.rs:20 */
    ZK_ERR_EVALUATE_ARITY = 2, /* "evaluate must assign to all variables"              evaluation_form.rs:85, product_poly.rs:38 */
    ZK_ERR_EMPTY_PRODUCT = 3,  /* "cannot create product polynomial from empty polynomials"   product_poly.rs:16 */
    ZK_ERR_NVARS_MISMATCH = 4, /* "cannot create product polynomial from polynomial that don't share the same number of variables" product_poly.rs:25 */
    ZK_ERR_PROOF_ROUNDS = 5,   /* "invalid proof: require 1 round poly for each variable in poly"   verifier.rs:18 */
    ZK_ERR_INITIAL_EVAL = 6,   /* "couldn't evaluate initial poly"                     verifier.rs:30 */
    ZK_ERR_ROUND_CHECK = 7,    /* "verifier check failed: claimed_sum != p(0) + p(1)"  verifier.rs:65 */
    ZK_VERIFY_FALSE = 8,       /* Ok(false): final oracle check failed                 verifier.rs:32 */
    ZK_ERR_NOT_POW2 = 9,       /* panic "values must be a power of 2"                  fft/src/lib.rs:29 */
    ZK_ERR_NO_ROOT = 10,       /* get_root_of_unity(..) == None -> unwrap panic        fft/src/lib.rs:6,14 */
    ZK_ERR_VAR_RANGE = 11,     /* initial_var + assignments run past n_vars (Rust: arithmetic-overflow panic, pairing_index.rs:3,6) */
    ZK_ERR_INVALID_ARG = 12,
    ZK_ERR_UNSUPPORTED = 13,
    ZK_ERR_CUDA = 14,
    ZK_ERR_NCCL = 15,
    ZK_ERR_OOM = 16
} zk_status;

ZK_API const char* zk_status_string(int status);
/* Message of the last non-OK status on this context (the reference's literal string, or CUDA/NCCL detail). */
ZK_API const char* zk_last_error(const zk_ctx* ctx);

/* ---- context ------------------------------------------------------------------------------ */
/* Single-GPU context on CUDA device `device`. */
ZK_API int zk_ctx_create(int device, zk_ctx** out);
/* Sharded context: this process is rank `rank` of `world` (power of two), one process per GPU.  Tables are
 * sharded by the LAST-bound variables (global index i lives on rank i mod world); round
 * polynomials are all-reduced over NCCL.  `nccl_id` = the 128 bytes produced by zk_nccl_unique_id on
 * rank 0 and distributed by the launcher (torch.distributed / MPI / a file). */
ZK_API int zk_nccl_unique_id(void* id_out_128);
ZK_API int zk_ctx_create_sharded(int device, int rank, int world, const void* nccl_id, zk_ctx** out);
/* Releases the context's streams, scratch, plans and communicator.  Tables created on it must be freed FIRST
 * (zk_table_free reads its context to select the device). */
ZK_API void zk_ctx_destroy(zk_ctx* ctx);
ZK_API int zk_ctx_rank(const zk_ctx* ctx);
ZK_API int zk_ctx_world(const zk_ctx* ctx);
/* 1 when the sharded context all-reduces the round sums through peer-mapped mailboxes inside the reducing launch (CUDA IPC
 * over NVLink; set up collectively at creation), 0 when it uses the NCCL all-reduce (IPC unavailable, ZK_B200_MAILBOX=0,
 * or a single rank).  The results are bit-identical either way. */
ZK_API int zk_ctx_uses_mailbox(const zk_ctx* ctx);
/* Local table length at or below which a sharded prove gathers the residual tables to every rank and
 * finishes without further collectives (default 4096). */
ZK_API int zk_ctx_set_gather_threshold(zk_ctx* ctx, uint64_t local_len);
/* Number of kernels this library launched on `ctx` since creation. */
ZK_API uint64_t zk_ctx_launch_count(const zk_ctx* ctx);
/* Device time (ms, CUDA events on the launching stream) of each round's kernel in the last
 * zk_sumcheck_prove* call; returns the number of rounds written (<= cap). */
ZK_API unsigned zk_ctx_last_round_ms(const zk_ctx* ctx, float* ms_out, unsigned cap);
/* Wall/device split of the last prove: [0] total ms (host clock), [1] ms in the initial-poly absorb,
 * [2] sum of round-kernel device ms. */
ZK_API int zk_ctx_last_prove_ms(const zk_ctx* ctx, double out[3]);
ZK_API int zk_ctx_synchronize(zk_ctx* ctx);
/* The cudaStream_t every kernel of this context is launched on (for callers that time with CUDA events). */
ZK_API void* zk_ctx_stream(const zk_ctx* ctx);

/* Pinned host memory for fast host<->device staging of large tables. */
ZK_API int zk_host_alloc(size_t bytes, void** out);
ZK_API int zk_host_free(void* p);

/* ---- MultiLinearPolynomial<F>  (polynomial/src/multilinear/evaluation_form.rs) ------------------- */
/* new(n_vars, evaluations) :15-27 — `len` must equal 2^n_vars else ZK_ERR_EVAL_LEN.  `mont_aos` is the
 * full table in host memory; on a sharded ctx each rank keeps entries rank, rank+world, ... */
ZK_API int zk_table_upload(zk_ctx* ctx, int field, const uint64_t* mont_aos, uint64_t len, unsigned n_vars, zk_table** out);
/* Deterministic synthetic table `table_id` (SURVEY.md 8d generator), generated on the device by global index. */
ZK_API int zk_table_generate(zk_ctx* ctx, int field, uint64_t seed, uint64_t table_id, unsigned n_vars, zk_table** out);
/* Refill an existing table (e.g. one consumed by zk_sumcheck_prove) with synthetic table `table_id`: no allocation. */
ZK_API int zk_table_regenerate(zk_ctx* ctx, zk_table* t, uint64_t seed, uint64_t table_id);
/* Upload this rank's local entries as they are (local_len * world must equal 2^n_vars): a shard the caller already
 * holds in the layout an entry point expects — the strided shard for the sumcheck entry points, the contiguous block
 * for the inverse zk_ntt_sharded. */
ZK_API int zk_table_upload_local(zk_ctx* ctx, int field, const uint64_t* local_mont_aos, uint64_t local_len, unsigned n_vars,
                                 zk_table** out);
ZK_API int zk_table_clone(zk_ctx* ctx, const zk_table* in, zk_table** out);
ZK_API void zk_table_free(zk_table* t);
ZK_API unsigned zk_table_n_vars(const zk_table* t);     /* n_vars() :30 */
ZK_API uint64_t zk_table_local_len(const zk_table* t);  /* entries held by this rank */
ZK_API int zk_table_field(const zk_table* t);
/* evaluation_slice() :92 — copies this rank's entries (all of them on a single-GPU ctx) to the host. */
ZK_API int zk_table_download(zk_ctx* ctx, const zk_table* t, uint64_t* mont_aos_out);
/* partial_evaluate(&self, initial_var, assignments) :40-80 — returns a NEW table with
 * n_vars - n_assign variables; variables initial_var .. initial_var+n_assign-1 are bound. */
ZK_API int zk_mle_partial_evaluate(zk_ctx* ctx, const zk_table* in, unsigned initial_var, const uint64_t* assignments,
                            unsigned n_assign, zk_table** out);
/* evaluate(&self, assignments) :83-89 — ZK_ERR_EVALUATE_ARITY unless len == n_vars. */
ZK_API int zk_mle_evaluate(zk_ctx* ctx, const zk_table* in, const uint64_t* point, unsigned len, uint64_t out[4]);
/* to_bytes() :97-103 — 32-byte big-endian canonical integers, 32 * 2^n_vars bytes. */
ZK_API int zk_mle_to_bytes(zk_ctx* ctx, const zk_table* in, uint8_t* out);

/* ---- ProductPoly<F>  (polynomial/src/product_poly.rs) ---------------------------------------- */
/* new(polynomials) :14-32 — the validation every product call performs: m == 0 -> ZK_ERR_EMPTY_PRODUCT,
 * differing n_vars -> ZK_ERR_NVARS_MISMATCH. */
ZK_API int zk_product_check(const zk_table* const* tables, unsigned m);
ZK_API int zk_product_evaluate(zk_ctx* ctx, const zk_table* const* tables, unsigned m, const uint64_t* point, unsigned len,
                        uint64_t out[4]);                                          /* evaluate :36-44 */
ZK_API int zk_product_prod_reduce(zk_ctx* ctx, const zk_table* const* tables, unsigned m, zk_table** out); /* :66-74 */
/* sum_j prod_k A_k[j] — `.prod_reduce().iter().sum()` on the unfolded product (the honest claim). */
ZK_API int zk_product_sum(zk_ctx* ctx, const zk_table* const* tables, unsigned m, uint64_t out[4]);
/* One round polynomial: evaluations at t = 0..degree of sum_x prod_k A_k(t, x)   (sumcheck/src/prover.rs:48-56). */
ZK_API int zk_product_round_poly(zk_ctx* ctx, const zk_table* const* tables, unsigned m, unsigned degree, uint64_t* out);
/* poly = poly.partial_evaluate(0, &[r]) in place (prover.rs:64): every table halves.  The in-place entry points (this
 * one, zk_product_fold_then_round_poly, zk_sumcheck_prove*) need DISTINCT tables: the same handle or device buffer listed
 * twice is ZK_ERR_INVALID_ARG — the reference's ProductPoly owns its factors (`vec![f.clone(), f.clone()]` are two
 * vectors), so pass a zk_table_clone for f * f (the Python / C++ / Rust mirrors do).  The read-only calls above accept
 * repeated handles. */
ZK_API int zk_product_fold_inplace(zk_ctx* ctx, zk_table* const* tables, unsigned m, const uint64_t r[4]);
/* The fused step: fold at r, then the next round polynomial, in one pass over the tables. */
ZK_API int zk_product_fold_then_round_poly(zk_ctx* ctx, zk_table* const* tables, unsigned m, unsigned degree,
                                    const uint64_t r[4], uint64_t* out);

/* ---- sumcheck  (sumcheck/src/prover.rs, verifier.rs, lib.rs) -------------------------------------- */
/* SumcheckProver::<degree,F>::prove (absorb_initial_poly != 0, prover.rs:15-20) or ::prove_partial
 * (== 0, prover.rs:24-30).  CONSUMES the tables (the reference takes `poly` by value): on return their
 * contents are unspecified and they may only be freed.  The tables must be distinct (see zk_product_fold_inplace).
 *   round_polys_out : n_vars * (degree+1) elements — SumcheckProof.round_polys (lib.rs:8-11)
 *   challenges_out  : n_vars elements (may be NULL)
 *   final_evals_out : m elements A_k(r_0..r_{n-1}) (may be NULL) */
ZK_API int zk_sumcheck_prove(zk_ctx* ctx, zk_table* const* tables, unsigned m, unsigned degree, const uint64_t sum[4],
                      int absorb_initial_poly, uint64_t* round_polys_out, uint64_t* challenges_out,
                      uint64_t* final_evals_out);
/* Same with HOST tables (each 2^n_vars elements): upload + (optional) claim + prove in one call — the
 * end-to-end entry a caller holding `Vec<Fr>`s uses.  If `sum` is NULL the claim is computed on the device
 * (zk_product_sum) and returned in sum_out.  On a sharded ctx each rank passes ITS shard (entries rank,
 * rank+world, ... stored densely, 2^n_vars / world elements) — no rank needs the whole table in host memory.
 * The device landing buffers (m tables) stay with the context between calls, grow-only, and are released by
 * zk_ctx_destroy: a proof per call then costs the uploads and the rounds, not a cudaMalloc/cudaFree of gigabytes.
 * Host tables should be pinned (zk_host_alloc); the uploads run on two copy streams. */
ZK_API int zk_sumcheck_prove_host(zk_ctx* ctx, int field, const uint64_t* const* host_tables, unsigned m, unsigned n_vars,
                           unsigned degree, const uint64_t* sum, int absorb_initial_poly, uint64_t* round_polys_out,
                           uint64_t* challenges_out, uint64_t* final_evals_out, uint64_t sum_out[4]);
/* SumcheckVerifier::verify (verifier.rs:15-33).  ZK_OK = Ok(true), ZK_VERIFY_FALSE = Ok(false),
 * ZK_ERR_PROOF_ROUNDS / ZK_ERR_ROUND_CHECK / ZK_ERR_INITIAL_EVAL = the Err cases.  Tables are not modified. */
ZK_API int zk_sumcheck_verify(zk_ctx* ctx, const zk_table* const* tables, unsigned m, const uint64_t sum[4],
                       const uint64_t* round_polys, unsigned n_rounds, unsigned degree);
/* SumcheckVerifier::verify_partial (verifier.rs:38-41): host only (ctx may be NULL).  Writes SubClaim
 * {sum, challenges} (lib.rs:17-20). */
ZK_API int zk_sumcheck_verify_partial(int field, const uint64_t sum[4], const uint64_t* round_polys, unsigned n_rounds,
                               unsigned degree, uint64_t subclaim_sum_out[4], uint64_t* challenges_out);

/* Proof dump for parity diffing (the reference defines no serialiser for SumcheckProof, lib.rs:8-11; layout of
 * SURVEY.md Appendix A.5): BE32(sum) || BE32(S_i(t)) for every round i, t = 0..degree || BE32(challenges) ||
 * BE32(final evaluations); the last two parts are omitted when their pointer is NULL.  Host only.  Writes the
 * byte count to *out_len (call with out == NULL to size the buffer) and, if digest_out != NULL, the Keccak-256
 * of the dump. */
ZK_API int zk_sumcheck_proof_dump(int field, const uint64_t sum[4], const uint64_t* round_polys, unsigned n_rounds,
                           unsigned degree, const uint64_t* challenges, const uint64_t* final_evals, unsigned m,
                           uint8_t* out, size_t out_cap, size_t* out_len, uint8_t digest_out[32]);

/* ---- sum of products ---------------------------------------------------------------
 * P(x) = sum_t prod_{k in term t} A_k(x) over n_tables distinct tables — the shape of a GKR layer polynomial
 * add.(Wb + Wc) + mul.Wb.Wc = add.Wb + add.Wc + mul.Wb.Wc, the caller readme.md:9 names.  The reference itself stops at a
 * single product (ProductPoly, polynomial/src/product_poly.rs:4-10), so there is no reference item behind these entry
 * points; the protocol is the reference's prover loop unchanged (sumcheck/src/prover.rs:33-73: append the sum, per round
 * the evaluations at t = 0..degree, a Keccak challenge, fold every table at it), and a single-term sum produces
 * bit-for-bit the ProductPoly proof.  Terms are given as `term_len[t]` factor counts (1..ZK_MAX_FACTORS) and the
 * concatenated `term_factors` (indices into tables[]); at most ZK_MAX_TERMS terms; a table may appear in several terms
 * or several times in one, but only once in tables[].  `degree` (MAX_VAR_DEGREE) must be 1..4 and, as in the
 * reference, is not validated against the longest term. */
#define ZK_MAX_TERMS 8
/* Evaluations at t = 0..degree of sum_x P(t, x): one round polynomial (prover.rs:48-56 for the sum of products). */
ZK_API int zk_sop_round_poly(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                             const uint8_t* term_factors, unsigned n_terms, unsigned degree, uint64_t* out);
/* sum over the hypercube of P: the honest claim. */
ZK_API int zk_sop_sum(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                      const uint8_t* term_factors, unsigned n_terms, uint64_t out[4]);
/* P(point): every table evaluated on the device (evaluation_form.rs:83-89), combined on the host — the verifier's
 * final check against SubClaim.sum (verifier.rs:28-32).  ZK_ERR_EVALUATE_ARITY unless len == n_vars. */
ZK_API int zk_sop_evaluate(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                           const uint8_t* term_factors, unsigned n_terms, const uint64_t* point, unsigned len,
                           uint64_t out[4]);
/* Host only: sum_t prod_k table_values[term_factors..] — P at a point from the n_tables table values there (e.g. the
 * prover's final evaluations). */
ZK_API int zk_sop_combine(int field, const uint8_t* term_len, const uint8_t* term_factors, unsigned n_terms,
                          const uint64_t* table_values, unsigned n_tables, uint64_t out[4]);
/* The prover over the sum of products; arguments as zk_sumcheck_prove (absorb_initial_poly absorbs the tables'
 * to_bytes() in tables[] order first).  CONSUMES the tables.  final_evals_out: n_tables elements.  Works on sharded
 * contexts like zk_sumcheck_prove (absorb_initial_poly == 0 there). */
ZK_API int zk_sumcheck_prove_sop(zk_ctx* ctx, zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                                 const uint8_t* term_factors, unsigned n_terms, unsigned degree, const uint64_t sum[4],
                                 int absorb_initial_poly, uint64_t* round_polys_out, uint64_t* challenges_out,
                                 uint64_t* final_evals_out);

/* SumcheckVerifier::verify (verifier.rs:15-33) for a proof made by zk_sumcheck_prove_sop with absorb_initial_poly != 0:
 * absorbs the tables' to_bytes(), replays the rounds, then checks the sub-claim against P(challenges) on the device.
 * Status codes as zk_sumcheck_verify.  Tables are not modified. */
ZK_API int zk_sumcheck_verify_sop(zk_ctx* ctx, const zk_table* const* tables, unsigned n_tables, const uint8_t* term_len,
                                  const uint8_t* term_factors, unsigned n_terms, const uint64_t sum[4],
                                  const uint64_t* round_polys, unsigned n_rounds, unsigned degree);

/* ---- transcript  (transcript/src/lib.rs) — host Keccak-256 -------------------------------------- */
ZK_API zk_transcript* zk_transcript_new(void);                                             /* :10 */
ZK_API void zk_transcript_free(zk_transcript* t);
ZK_API void zk_transcript_append(zk_transcript* t, const uint8_t* data, size_t len);        /* :16 */
ZK_API int zk_transcript_sample_field_element(zk_transcript* t, int field, uint64_t out[4]); /* :27 */
ZK_API int zk_transcript_sample_n_field_elements(zk_transcript* t, int field, unsigned n, uint64_t* out); /* :32 */
ZK_API void zk_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);

/* ---- fft  (fft/src/lib.rs) ------------------------------------------------------------------ */
/* fft (:4-8) / ifft (:11-19) of the 2^n_vars-entry table, in place, natural order in and out.
 * ZK_ERR_NO_ROOT if 2^n_vars exceeds the field's two-adic subgroup. */
ZK_API int zk_ntt(zk_ctx* ctx, zk_table* inout, int inverse);
/* Host-buffer form: `len` must be a power of two (else ZK_ERR_NOT_POW2). */
ZK_API int zk_ntt_host(zk_ctx* ctx, int field, uint64_t* mont_aos_inout, uint64_t len, int inverse);
/* Multi-GPU fft / ifft on a sharded context (collective: every rank calls it).  Layout contract:
 *   forward : `inout` holds this rank's STRIDED shard a[rank], a[rank + world], ... (what zk_table_upload /
 *             zk_table_generate keep on a sharded context); on return it holds the CONTIGUOUS block
 *             X[rank * N/world .. (rank+1) * N/world) of the natural-order transform;
 *   inverse : contiguous block in (zk_table_upload_local), strided shard out — ifft(fft(a)) round-trips in place.
 * world in {2, 4, 8}; N >= world^2 (else ZK_ERR_UNSUPPORTED); ZK_ERR_NO_ROOT as for zk_ntt.  Per rank: the
 * single-GPU transform of its N/world entries, one twiddle multiplication per entry, two all-to-all exchanges (NCCL
 * send/recv groups over NVLink) around a world-point DFT.  On an unsharded context it is zk_ntt. */
ZK_API int zk_ntt_sharded(zk_ctx* ctx, zk_table* inout, int inverse);
/* The same factorisation with `ranks` VIRTUAL ranks on one GPU (unsharded ctx, full table in and out, natural order,
 * result identical to zk_ntt): every kernel and index map of the multi-GPU path with device-to-device copies in
 * place of the NCCL exchanges — how the path is validated on a single GPU. */
ZK_API int zk_ntt_virtual_sharded(zk_ctx* ctx, zk_table* inout, unsigned ranks, int inverse);

/* ---- field helpers for harnesses (host; ark-ff conversions the reference calls) --------------------- */
ZK_API int zk_field_from_canonical(int field, const uint64_t* canon, uint64_t* mont, size_t count); /* reduces mod p */
ZK_API int zk_field_to_canonical(int field, const uint64_t* mont, uint64_t* canon, size_t count);   /* into_bigint */
ZK_API int zk_field_from_u64(int field, uint64_t x, uint64_t out[4]);                               /* F::from(u64) */
ZK_API int zk_field_to_bytes_be(int field, const uint64_t* mont, size_t count, uint8_t* out);       /* to_bytes_be */
ZK_API int zk_field_from_be_bytes_mod_order(int field, const uint8_t in[32], uint64_t out[4]);
ZK_API int zk_field_mul(int field, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
ZK_API int zk_field_add(int field, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
ZK_API int zk_field_sub(int field, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
ZK_API int zk_field_inverse(int field, const uint64_t a[4], uint64_t out[4]); /* ZK_ERR_INVALID_ARG for zero */
ZK_API int zk_field_root_of_unity(int field, uint64_t n, uint64_t out[4]);    /* get_root_of_unity; ZK_ERR_NO_ROOT */
/* Value at x of the polynomial of degree < n_points given by its evaluations at 0..n_points-1:
 * `UnivariatePolynomial::interpolate(xs = 0.., ys).evaluate(x)` (polynomial/src/univariate_poly.rs:29-80), what the
 * verifier computes as the next claimed sum (sumcheck/src/verifier.rs:68-70) and what the prover uses to derive
 * S_i(1) = S_{i-1}(r_{i-1}) - S_i(0) in rounds >= 1.  Host only; n_points in 1..ZK_MAX_DEGREE+1. */
ZK_API int zk_round_poly_evaluate(int field, const uint64_t* ys, unsigned n_points, const uint64_t x[4], uint64_t out[4]);

/* ---- measurement support ---------------------------------------------------------------------- */
typedef struct zk_microbench {
    double imad_wide_per_s; /* IMAD.WIDE.U32 issued per second, whole chip */
    double imad_lo_per_s;
    double iadd3_per_s;
    double mixed_per_s;
    double fe_mul_per_s; /* stand-alone Montgomery multiplications per second */
    double copy_gbs;     /* 256-bit streaming copy, read+write GB/s */
    double read_gbs;
    double sm_clock_mhz;
    double dfma_per_s;       /* FP64 FMA issued per second, whole chip (the folds run on this pipe) */
    double fe_mul_fixed_per_s; /* stand-alone fixed-multiplier products per second, FP64-pipe version */
} zk_microbench;
ZK_API int zk_microbench_run(zk_ctx* ctx, int field, zk_microbench* out);

#ifdef __cplusplus
}
#endif
#endif /* ZK_B200_H */
