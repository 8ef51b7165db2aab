/*
 * oracle/cpu_ref.c — CPU oracle (plain C restatement) of the iammadab/zk sumcheck / MLE / FFT path.
 *
 * TEST INFRASTRUCTURE ONLY.  The product (zk_b200/, libzk_b200.so) never links, loads or calls
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do, and there only as the checker / the reported CPU baseline.
 *
 * The Rust reference cannot be compiled here (no rustc/cargo; field arithmetic lives in the
 * un-vendored crates ark-ff 0.5.0, ark-bls12-381 0.5.0, ark-bls12-377 0.5.0; hashing in sha3 0.10.8),
 * so oracle/_ref does not exist.  This file restates the reference's algorithm with the same
 * *shape* (single thread, per-call table clones, separate prod_reduce and sum passes, one
 * Montgomery multiplication per field multiplication, 4x64-bit limbs like ark-ff's
 * Fp<MontBackend<_,4>,4>), each function citing the reference file:line it follows (paths
 * relative to /root/reference).  Pinning: checked against oracle/zkoracle.py (big-int) and every
 * reference KAT for the path in tests/test_oracle_kats.py.  Values the reference never asserts
 * (transcript bytes, round polynomials, forward-NTT values) are pinned only by Keccak-256 KATs,
 * a naive DFT and SURVEY.md Appendix B — "parity unpinned by the reference" for those.
 *
 * Element layout everywhere: 4 little-endian uint64 limbs, Montgomery form (x * 2^256 mod p),
 * fully reduced — the in-memory layout of ark-ff 0.5 `Fp256`.
 */
#include <stdint.h>
#include <stdlib.h>
#include <pthread.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;

typedef struct {
    uint64_t p[4];
    uint64_t r[4];    /* R   = 2^256 mod p  (Montgomery one) */
    uint64_t r2[4];   /* R^2 mod p */
    uint64_t inv;     /* -p^-1 mod 2^64 */
    uint64_t gen;     /* multiplicative generator (small) */
    unsigned two_adicity;
} field_t;

static const field_t FIELDS[2] = {
    /* ark-bls12-381 0.5.0 Fr */
    {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
     {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL},
     {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL},
     0xfffffffeffffffffULL, 7, 32},
    /* ark-bls12-377 0.5.0 Fr */
    {{0x0a11800000000001ULL, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL, 0x12ab655e9a2ca556ULL},
     {0x7d1c7ffffffffff3ULL, 0x7257f50f6ffffff2ULL, 0x16d81575512c0feeULL, 0x0d4bda322bbb9a9dULL},
     {0x25d577bab861857bULL, 0xcc2c27b58860591fULL, 0xa7cc008fe5dc8593ULL, 0x011fdae7eff1c939ULL},
     0x0a117fffffffffffULL, 22, 47},
};

/* ------------------------------------------------------------------------------------------
 * Field arithmetic (ark-ff semantics: exact ops on fully reduced Montgomery residues)
 * ---------------------------------------------------------------------------------------- */
static inline int ge_p(const uint64_t a[4], const field_t *F) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > F->p[i]) return 1;
        if (a[i] < F->p[i]) return 0;
    }
    return 1;
}
static inline void sub_p(uint64_t a[4], const field_t *F) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - F->p[i] - (uint64_t)b;
        a[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
}
static inline fe f_add(const fe *a, const fe *b, const field_t *F) {
    fe r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    if (c || ge_p(r.v, F)) sub_p(r.v, F);
    return r;
}
static inline fe f_sub(const fe *a, const fe *b, const field_t *F) {
    fe r; u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->v[i] - b->v[i] - (uint64_t)br;
        r.v[i] = (uint64_t)d; br = (d >> 64) & 1;
    }
    if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + F->p[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
    return r;
}
/* CIOS Montgomery multiplication, 4x64 limbs */
static inline fe f_mul(const fe *a, const fe *b, const field_t *F) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = ((u128)m * F->p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * F->p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe r; memcpy(r.v, t, 32);
    if (t[4] || ge_p(r.v, F)) sub_p(r.v, F);
    return r;
}
static inline int f_is_zero(const fe *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int f_eq(const fe *a, const fe *b) { return memcmp(a->v, b->v, 32) == 0; }
static inline fe f_zero(void) { fe z = {{0, 0, 0, 0}}; return z; }
static inline fe f_one(const field_t *F) { fe o; memcpy(o.v, F->r, 32); return o; }
static inline fe f_from_canonical(const uint64_t c[4], const field_t *F) {  /* c < p */
    fe a, r2; memcpy(a.v, c, 32); memcpy(r2.v, F->r2, 32); return f_mul(&a, &r2, F);
}
static inline fe f_from_u64(uint64_t x, const field_t *F) {                  /* F::from(u64) */
    uint64_t c[4] = {x, 0, 0, 0};
    return f_from_canonical(c, F);
}
static inline void f_to_canonical(const fe *a, uint64_t out[4], const field_t *F) {  /* into_bigint */
    fe one = {{1, 0, 0, 0}}; fe r = f_mul(a, &one, F); memcpy(out, r.v, 32);
}
static fe f_pow(const fe *a, const uint64_t *e, int nlimbs, const field_t *F) {
    fe r = f_one(F);
    for (int i = nlimbs - 1; i >= 0; i--)
        for (int b = 63; b >= 0; b--) { r = f_mul(&r, &r, F); if ((e[i] >> b) & 1) r = f_mul(&r, a, F); }
    return r;
}
static fe f_inv(const fe *a, const field_t *F) {   /* a^(p-2); caller guarantees a != 0 */
    uint64_t e[4]; memcpy(e, F->p, 32); e[0] -= 2; /* p[0] >= 2 for both fields */
    return f_pow(a, e, 4, F);
}
static void f_to_be32(const fe *a, uint8_t out[32], const field_t *F) {  /* into_bigint().to_bytes_be() */
    uint64_t c[4]; f_to_canonical(a, c, F);
    for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) out[31 - (8 * i + b)] = (uint8_t)(c[i] >> (8 * b));
}
/* F::from_be_bytes_mod_order(32 bytes): value < 2^256 reduced mod p.  2^256 < 3p (381) / < 14p (377). */
static fe f_from_be32_mod_order(const uint8_t in[32], const field_t *F) {
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) c[i] |= (uint64_t)in[31 - (8 * i + b)] << (8 * b);
    while (ge_p(c, F)) sub_p(c, F);
    return f_from_canonical(c, F);
}

/* ------------------------------------------------------------------------------------------
 * Keccak-256 (sha3 0.10.8 Keccak256: rate 136, pad 0x01 .. 0x80) and the Transcript
 * (transcript/src/lib.rs:5-35)
 * ---------------------------------------------------------------------------------------- */
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#define ROL64(x, n) ((n) ? (((x) << (n)) | ((x) >> (64 - (n)))) : (x))
static void keccak_f(uint64_t s[25]) {   /* s[x + 5y] */
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ ROL64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) s[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) {
            int i = x + 5 * y; int nx = y, ny = (2 * x + 3 * y) % 5;
            B[nx + 5 * ny] = ROL64(s[i], KROT[i]);
        }
        for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++)
            s[x + 5 * y] = B[x + 5 * y] ^ ((~B[(x + 1) % 5 + 5 * y]) & B[(x + 2) % 5 + 5 * y]);
        s[0] ^= KRC[rnd];
    }
}
typedef struct { uint64_t s[25]; uint8_t buf[136]; unsigned fill; } keccak_t;
static void k_init(keccak_t *k) { memset(k, 0, sizeof *k); }
static void k_block(keccak_t *k, const uint8_t *b) {
    for (int i = 0; i < 17; i++) { uint64_t l; memcpy(&l, b + 8 * i, 8); k->s[i] ^= l; }  /* little-endian host */
    keccak_f(k->s);
}
static void k_update(keccak_t *k, const uint8_t *d, size_t n) {
    if (k->fill) {
        size_t take = 136 - k->fill; if (take > n) take = n;
        memcpy(k->buf + k->fill, d, take); k->fill += (unsigned)take; d += take; n -= take;
        if (k->fill == 136) { k_block(k, k->buf); k->fill = 0; }
    }
    while (n >= 136) { k_block(k, d); d += 136; n -= 136; }
    if (n) { memcpy(k->buf, d, n); k->fill = (unsigned)n; }
}
static void k_finalize_reset(keccak_t *k, uint8_t out[32]) {
    memset(k->buf + k->fill, 0, 136 - k->fill);
    k->buf[k->fill] ^= 0x01; k->buf[135] ^= 0x80;
    k_block(k, k->buf);
    memcpy(out, k->s, 32);
    k_init(k);
}
/* Transcript::sample_field_element — transcript/src/lib.rs:20-30 */
static fe t_sample(keccak_t *k, const field_t *F) {
    uint8_t d[32]; k_finalize_reset(k, d); k_update(k, d, 32);
    return f_from_be32_mod_order(d, F);
}

/* ------------------------------------------------------------------------------------------
 * MLE partial evaluation — polynomial/src/multilinear/evaluation_form.rs:40-80
 * (pair addressing: polynomial/src/multilinear/pairing_index.rs:2-21)
 * Returns a freshly malloc'd table of 2^(n_vars - n_assign) entries, like the reference's
 * clone (:49) + to_vec (:78).
 * ---------------------------------------------------------------------------------------- */
static fe *mle_partial_evaluate(const fe *evals, unsigned n_vars, unsigned initial_var,
                                const fe *assign, unsigned n_assign, const field_t *F) {
    size_t n = (size_t)1 << n_vars;
    fe *w = (fe *)malloc(n * sizeof(fe));
    memcpy(w, evals, n * sizeof(fe));                                  /* :49 clone */
    fe one = f_one(F);
    for (unsigned s = 0; s < n_assign; s++) {                          /* :54 */
        unsigned nv = n_vars - s, pos = nv - 1 - initial_var;          /* index_pair(nv, initial_var) */
        size_t pairs = (size_t)1 << (nv - 1), low_mask = ((size_t)1 << pos) - 1;
        int a0 = f_is_zero(&assign[s]), a1 = f_eq(&assign[s], &one);
        for (size_t k = 0; k < pairs; k++) {                           /* :56 */
            size_t l = ((k >> pos) << (pos + 1)) | (k & low_mask);      /* insert_bit(k, pos, 0) */
            size_t r = l | ((size_t)1 << pos);
            fe left = w[l], right = w[r];
            if (a0) w[k] = left;                                       /* :61 */
            else if (a1) w[k] = right;                                 /* :62 */
            else {                                                     /* :68 left - a*(left-right) */
                fe d = f_sub(&left, &right, F); fe m = f_mul(&assign[s], &d, F); w[k] = f_sub(&left, &m, F);
            }
        }
    }
    size_t out_n = (size_t)1 << (n_vars - n_assign);
    fe *out = (fe *)malloc(out_n * sizeof(fe));                        /* :78 to_vec */
    memcpy(out, w, out_n * sizeof(fe));
    free(w);
    return out;
}

/* ==========================================================================================
 * Exported C surface (ctypes): names prefixed zko_.  All element buffers are uint64[4*count].
 * ======================================================================================== */
#define EXPORT __attribute__((visibility("default")))

EXPORT void zko_keccak256(const uint8_t *data, size_t n, uint8_t out[32]) {
    keccak_t k; k_init(&k); k_update(&k, data, n); k_finalize_reset(&k, out);
}
EXPORT void zko_to_mont(int field, const uint64_t *canon, uint64_t *mont, size_t count) {
    const field_t *F = &FIELDS[field];
    for (size_t i = 0; i < count; i++) { fe r = f_from_canonical(canon + 4 * i, F); memcpy(mont + 4 * i, r.v, 32); }
}
EXPORT void zko_from_mont(int field, const uint64_t *mont, uint64_t *canon, size_t count) {
    const field_t *F = &FIELDS[field];
    for (size_t i = 0; i < count; i++) { fe a; memcpy(a.v, mont + 4 * i, 32); f_to_canonical(&a, canon + 4 * i, F); }
}
EXPORT void zko_mul(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count) {
    const field_t *F = &FIELDS[field];
    for (size_t i = 0; i < count; i++) {
        fe x, y; memcpy(x.v, a + 4 * i, 32); memcpy(y.v, b + 4 * i, 32);
        fe r = f_mul(&x, &y, F); memcpy(out + 4 * i, r.v, 32);
    }
}

/* Synthetic table generator: counter-based splitmix64 keyed by GLOBAL index. */
static inline uint64_t splitmix64_mix(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31; return z;
}
EXPORT void zko_gen_table(int field, uint64_t seed, uint64_t table_id, unsigned n_vars,
                          uint64_t first, uint64_t stride, uint64_t count, uint64_t *out_mont) {
    /* entries first, first+stride, ... (count of them) of the 2^n_vars-entry table */
    const field_t *F = &FIELDS[field]; (void)n_vars;
    for (uint64_t j = 0; j < count; j++) {
        uint64_t i = first + j * stride, c[4];
        for (int l = 0; l < 4; l++) {
            uint64_t ctr = ((table_id << 40) + i) * 4 + (uint64_t)l;
            c[l] = splitmix64_mix(seed + 0x9E3779B97F4A7C15ULL * (ctr + 1));
        }
        c[3] &= 0x3FFFFFFFFFFFFFFFULL;
        while (ge_p(c, F)) sub_p(c, F);     /* only bites for the 253-bit field */
        fe r = f_from_canonical(c, F); memcpy(out_mont + 4 * j, r.v, 32);
    }
}

EXPORT void zko_partial_evaluate(int field, const uint64_t *evals, unsigned n_vars, unsigned initial_var,
                                 const uint64_t *assign, unsigned n_assign, uint64_t *out) {
    const field_t *F = &FIELDS[field];
    fe *r = mle_partial_evaluate((const fe *)evals, n_vars, initial_var, (const fe *)assign, n_assign, F);
    memcpy(out, r, ((size_t)1 << (n_vars - n_assign)) * sizeof(fe));
    free(r);
}

/* sum_j prod_k A_k[j] — the claim (ProductPoly::prod_reduce, product_poly.rs:66-74, then .sum()) */
EXPORT void zko_product_sum(int field, const uint64_t *const *tables, unsigned m, unsigned n_vars, uint64_t out[4]) {
    const field_t *F = &FIELDS[field];
    size_t n = (size_t)1 << n_vars; fe acc = f_zero();
    for (size_t j = 0; j < n; j++) {
        fe pr = ((const fe *)tables[0])[j];
        for (unsigned k = 1; k < m; k++) pr = f_mul(&pr, &((const fe *)tables[k])[j], F);
        acc = f_add(&acc, &pr, F);
    }
    memcpy(out, acc.v, 32);
}

/* ------------------------------------------------------------------------------------------
 * SumcheckProver::prove / prove_partial — sumcheck/src/prover.rs:15-73, REFERENCE-SHAPED:
 * per evaluation point t a full partial_evaluate (clone + fold) of every factor, a materialised
 * prod_reduce vector, then a separate sum pass; then the challenge fold.  This is the timed
 * CPU baseline ("kind": "port").
 *   round_polys_out: n_vars*(degree+1) elements, challenges_out: n_vars elements,
 *   finals_out: m elements (the fully folded factors; may be NULL).
 * absorb != 0 => `prove` (absorbs poly.to_bytes() first, :16-17), else `prove_partial`.
 * ---------------------------------------------------------------------------------------- */
EXPORT int zko_prove(int field, const uint64_t *const *tables, unsigned m, unsigned n_vars, unsigned degree,
                     const uint64_t sum[4], int absorb, uint64_t *round_polys_out, uint64_t *challenges_out,
                     uint64_t *finals_out) {
    const field_t *F = &FIELDS[field];
    keccak_t tr; k_init(&tr);                                           /* Transcript::new */
    size_t n = (size_t)1 << n_vars;
    fe **poly = (fe **)malloc(m * sizeof(fe *));
    for (unsigned k = 0; k < m; k++) { poly[k] = (fe *)malloc(n * sizeof(fe)); memcpy(poly[k], tables[k], n * sizeof(fe)); }
    uint8_t be[32];
    if (absorb) {                                                       /* poly.to_bytes(): product_poly.rs:77-83 */
        for (unsigned k = 0; k < m; k++) for (size_t j = 0; j < n; j++) { f_to_be32(&poly[k][j], be, F); k_update(&tr, be, 32); }
    }
    fe s; memcpy(s.v, sum, 32); f_to_be32(&s, be, F); k_update(&tr, be, 32);   /* :42 */
    unsigned nv = n_vars;
    for (unsigned round = 0; round < n_vars; round++) {                 /* :44 */
        size_t half = (size_t)1 << (nv - 1);
        for (unsigned t = 0; t <= degree; t++) {                        /* :49 */
            fe ft = f_from_u64(t, F);
            fe **pe = (fe **)malloc(m * sizeof(fe *));
            for (unsigned k = 0; k < m; k++) pe[k] = mle_partial_evaluate(poly[k], nv, 0, &ft, 1, F);   /* :51 */
            fe *prod = (fe *)malloc(half * sizeof(fe));                 /* prod_reduce: product_poly.rs:66-74 */
            memcpy(prod, pe[0], half * sizeof(fe));
            for (unsigned k = 1; k < m; k++) for (size_t j = 0; j < half; j++) prod[j] = f_mul(&prod[j], &pe[k][j], F);
            fe acc = f_zero();
            for (size_t j = 0; j < half; j++) acc = f_add(&acc, &prod[j], F);   /* :53-54 */
            memcpy(round_polys_out + 4 * ((size_t)round * (degree + 1) + t), acc.v, 32);
            f_to_be32(&acc, be, F); k_update(&tr, be, 32);              /* :59 (appended in order t=0..D) */
            free(prod); for (unsigned k = 0; k < m; k++) free(pe[k]); free(pe);
        }
        fe r = t_sample(&tr, F);                                        /* :62 */
        memcpy(challenges_out + 4 * (size_t)round, r.v, 32);
        for (unsigned k = 0; k < m; k++) {                              /* :64 */
            fe *nx = mle_partial_evaluate(poly[k], nv, 0, &r, 1, F); free(poly[k]); poly[k] = nx;
        }
        nv--;
    }
    for (unsigned k = 0; k < m; k++) { if (finals_out) memcpy(finals_out + 4 * k, poly[k][0].v, 32); free(poly[k]); }
    free(poly);
    return 0;
}

/* Streamlined variant (same results, fused single pass per round, in place) for cross-checks at
 * sizes where the reference-shaped prover is too slow.  NOT the reported baseline. */
EXPORT int zko_prove_fast(int field, uint64_t *const *tables, unsigned m, unsigned n_vars, unsigned degree,
                          const uint64_t sum[4], int absorb, uint64_t *round_polys_out, uint64_t *challenges_out,
                          uint64_t *finals_out) {
    const field_t *F = &FIELDS[field];
    if (m > 8 || degree > 15) return -1;
    keccak_t tr; k_init(&tr);
    uint8_t be[32];
    if (absorb) {                                                       /* `prove`: poly.to_bytes() first (prover.rs:16-17) */
        const size_t n = (size_t)1 << n_vars;
        for (unsigned k = 0; k < m; k++) for (size_t j = 0; j < n; j++) { f_to_be32(&((const fe *)tables[k])[j], be, F); k_update(&tr, be, 32); }
    }
    fe s; memcpy(s.v, sum, 32); f_to_be32(&s, be, F); k_update(&tr, be, 32);
    unsigned nv = n_vars;
    for (unsigned round = 0; round < n_vars; round++) {
        size_t half = (size_t)1 << (nv - 1);
        fe acc[16]; for (unsigned t = 0; t <= degree; t++) acc[t] = f_zero();
        for (size_t j = 0; j < half; j++) {
            fe e[8], d[8];
            for (unsigned k = 0; k < m; k++) { fe *T = (fe *)tables[k]; e[k] = T[j]; d[k] = f_sub(&T[j + half], &T[j], F); }
            for (unsigned t = 0; t <= degree; t++) {
                fe pr = e[0];
                for (unsigned k = 1; k < m; k++) pr = f_mul(&pr, &e[k], F);
                acc[t] = f_add(&acc[t], &pr, F);
                for (unsigned k = 0; k < m; k++) e[k] = f_add(&e[k], &d[k], F);
            }
        }
        for (unsigned t = 0; t <= degree; t++) {
            memcpy(round_polys_out + 4 * ((size_t)round * (degree + 1) + t), acc[t].v, 32);
            f_to_be32(&acc[t], be, F); k_update(&tr, be, 32);
        }
        fe r = t_sample(&tr, F);
        memcpy(challenges_out + 4 * (size_t)round, r.v, 32);
        for (unsigned k = 0; k < m; k++) {
            fe *T = (fe *)tables[k];
            for (size_t j = 0; j < half; j++) { fe d = f_sub(&T[j], &T[j + half], F); fe x = f_mul(&r, &d, F); T[j] = f_sub(&T[j], &x, F); }
        }
        nv--;
    }
    if (finals_out) for (unsigned k = 0; k < m; k++) memcpy(finals_out + 4 * k, tables[k], 32);
    return 0;
}

/* The streamlined prover on several host cores (pthreads over the pairs of a round; modular sums are order
 * independent, so the proof is bit-identical).  Reported by bench.py NEXT TO the single-threaded reference-shaped number
 * as "what a multi-core CPU port of the same algorithm does" — the reference itself is single-threaded. */
typedef struct {
    const field_t *F; uint64_t *const *tables; unsigned m, degree; long half, lo, hi; fe r; fe loc[16];
} mt_job;
static void *mt_sum_worker(void *arg) {
    mt_job *J = (mt_job *)arg; const field_t *F = J->F;
    for (unsigned t = 0; t <= J->degree; t++) J->loc[t] = f_zero();
    for (long j = J->lo; j < J->hi; j++) {
        fe e[8], d[8];
        for (unsigned k = 0; k < J->m; k++) { fe *T = (fe *)J->tables[k]; e[k] = T[j]; d[k] = f_sub(&T[j + J->half], &T[j], F); }
        for (unsigned t = 0; t <= J->degree; t++) {
            fe pr = e[0];
            for (unsigned k = 1; k < J->m; k++) pr = f_mul(&pr, &e[k], F);
            J->loc[t] = f_add(&J->loc[t], &pr, F);
            for (unsigned k = 0; k < J->m; k++) e[k] = f_add(&e[k], &d[k], F);
        }
    }
    return NULL;
}
static void *mt_fold_worker(void *arg) {
    mt_job *J = (mt_job *)arg; const field_t *F = J->F;
    for (unsigned k = 0; k < J->m; k++) {
        fe *T = (fe *)J->tables[k];
        for (long j = J->lo; j < J->hi; j++) { fe d = f_sub(&T[j], &T[j + J->half], F); fe x = f_mul(&J->r, &d, F); T[j] = f_sub(&T[j], &x, F); }
    }
    return NULL;
}
static void mt_run(void *(*fn)(void *), mt_job *jobs, int n_jobs) {
    pthread_t th[64];
    for (int i = 1; i < n_jobs; i++) pthread_create(&th[i], NULL, fn, &jobs[i]);
    fn(&jobs[0]);
    for (int i = 1; i < n_jobs; i++) pthread_join(th[i], NULL);
}
/* rounds first_round .. n_vars-1 of the streamlined prover on tables of 2^(n_vars-first_round) entries, continuing
 * the transcript `tr` */
static void mt_rounds(const field_t *F, keccak_t *tr, uint64_t *const *tables, unsigned m, unsigned n_vars, unsigned first_round,
                      unsigned degree, uint64_t *round_polys_out, uint64_t *challenges_out, int n_threads) {
    uint8_t be[32];
    unsigned nv = n_vars - first_round;
    mt_job jobs[64];
    for (unsigned round = first_round; round < n_vars; round++) {
        const long half = (long)1 << (nv - 1);
        const int nj = half >= 4096 ? n_threads : 1;   /* small rounds: thread start-up would dominate */
        for (int i = 0; i < nj; i++) {
            jobs[i].F = F; jobs[i].tables = tables; jobs[i].m = m; jobs[i].degree = degree; jobs[i].half = half;
            jobs[i].lo = half * i / nj; jobs[i].hi = half * (i + 1) / nj;
        }
        mt_run(mt_sum_worker, jobs, nj);
        fe acc[16];
        for (unsigned t = 0; t <= degree; t++) {
            acc[t] = f_zero();
            for (int i = 0; i < nj; i++) acc[t] = f_add(&acc[t], &jobs[i].loc[t], F);
            memcpy(round_polys_out + 4 * ((size_t)round * (degree + 1) + t), acc[t].v, 32);
            f_to_be32(&acc[t], be, F); k_update(tr, be, 32);
        }
        fe r = t_sample(tr, F);
        memcpy(challenges_out + 4 * (size_t)round, r.v, 32);
        for (int i = 0; i < nj; i++) jobs[i].r = r;
        mt_run(mt_fold_worker, jobs, nj);
        nv--;
    }
}
EXPORT int zko_prove_fast_mt(int field, uint64_t *const *tables, unsigned m, unsigned n_vars, unsigned degree,
                             const uint64_t sum[4], uint64_t *round_polys_out, uint64_t *challenges_out,
                             uint64_t *finals_out, int n_threads) {
    const field_t *F = &FIELDS[field];
    if (m > 8 || degree > 15 || n_threads < 1 || n_threads > 64) return -1;
    keccak_t tr; k_init(&tr);
    uint8_t be[32]; fe s; memcpy(s.v, sum, 32); f_to_be32(&s, be, F); k_update(&tr, be, 32);
    mt_rounds(F, &tr, tables, m, n_vars, 0, degree, round_polys_out, challenges_out, n_threads);
    if (finals_out) for (unsigned k = 0; k < m; k++) memcpy(finals_out + 4 * k, tables[k], 32);
    return 0;
}

/* The same proof for the SEEDED synthetic tables (zko_gen_table, table ids 0..m-1) at sizes whose tables do not fit the
 * host: round 0 streams the generator (no table is ever materialised at 2^n_vars entries), the fold at the first
 * challenge regenerates every pair and writes the 2^(n_vars-1)-entry tables, the remaining rounds are mt_rounds.
 * The claimed sum is the true sum S_0(0) + S_0(1) (returned in claim_out) — exactly what
 * `prove_partial(poly, poly.sum())` absorbs (prover.rs:42).  Used offline by tests/golden/make_fullsize_digests.py for
 * 2^27 .. 2^30 entries; the CPU suite checks it against zko_prove at small sizes. */
typedef struct {
    const field_t *F; int field; uint64_t seed; unsigned m, degree, n_vars; long half, lo, hi; fe r; fe loc[16]; fe **out;
} gen_job;
static void gen_pair(const gen_job *J, unsigned k, long j, fe *lo, fe *hi) {
    zko_gen_table(J->field, J->seed, k, J->n_vars, (uint64_t)j, 1, 1, lo->v);
    zko_gen_table(J->field, J->seed, k, J->n_vars, (uint64_t)(j + J->half), 1, 1, hi->v);
}
static void *gen_sum_worker(void *arg) {
    gen_job *J = (gen_job *)arg; const field_t *F = J->F;
    for (unsigned t = 0; t <= J->degree; t++) J->loc[t] = f_zero();
    for (long j = J->lo; j < J->hi; j++) {
        fe e[8], d[8], hi;
        for (unsigned k = 0; k < J->m; k++) { gen_pair(J, k, j, &e[k], &hi); d[k] = f_sub(&hi, &e[k], F); }
        for (unsigned t = 0; t <= J->degree; t++) {
            fe pr = e[0];
            for (unsigned k = 1; k < J->m; k++) pr = f_mul(&pr, &e[k], F);
            J->loc[t] = f_add(&J->loc[t], &pr, F);
            for (unsigned k = 0; k < J->m; k++) e[k] = f_add(&e[k], &d[k], F);
        }
    }
    return NULL;
}
static void *gen_fold_worker(void *arg) {
    gen_job *J = (gen_job *)arg; const field_t *F = J->F;
    for (unsigned k = 0; k < J->m; k++)
        for (long j = J->lo; j < J->hi; j++) {
            fe lo, hi; gen_pair(J, k, j, &lo, &hi);
            fe d = f_sub(&lo, &hi, F); fe x = f_mul(&J->r, &d, F); J->out[k][j] = f_sub(&lo, &x, F);   /* evaluation_form.rs:68 */
        }
    return NULL;
}
static void gen_run(void *(*fn)(void *), gen_job *jobs, int n_jobs) {
    pthread_t th[64];
    for (int i = 1; i < n_jobs; i++) pthread_create(&th[i], NULL, fn, &jobs[i]);
    fn(&jobs[0]);
    for (int i = 1; i < n_jobs; i++) pthread_join(th[i], NULL);
}
EXPORT int zko_prove_generated_mt(int field, uint64_t seed, unsigned m, unsigned n_vars, unsigned degree, uint64_t claim_out[4],
                                  uint64_t *round_polys_out, uint64_t *challenges_out, uint64_t *finals_out, int n_threads) {
    const field_t *F = &FIELDS[field];
    if (m > 8 || degree > 15 || degree < 1 || n_threads < 1 || n_threads > 64 || n_vars < 1) return -1;
    const long half = (long)1 << (n_vars - 1);
    gen_job jobs[64];
    const int nj = half >= 64 ? n_threads : 1;
    fe *tabs[8];
    for (unsigned k = 0; k < m; k++) { tabs[k] = (fe *)malloc((size_t)half * sizeof(fe)); if (!tabs[k]) return -2; }
    for (int i = 0; i < nj; i++) {
        jobs[i].F = F; jobs[i].field = field; jobs[i].seed = seed; jobs[i].m = m; jobs[i].degree = degree; jobs[i].n_vars = n_vars;
        jobs[i].half = half; jobs[i].lo = half * i / nj; jobs[i].hi = half * (i + 1) / nj; jobs[i].out = tabs;
    }
    gen_run(gen_sum_worker, jobs, nj);
    fe acc[16];
    for (unsigned t = 0; t <= degree; t++) {
        acc[t] = f_zero();
        for (int i = 0; i < nj; i++) acc[t] = f_add(&acc[t], &jobs[i].loc[t], F);
    }
    keccak_t tr; k_init(&tr);
    uint8_t be[32];
    fe claim = f_add(&acc[0], &acc[1], F);                              /* the true sum: sum over the last variable first */
    memcpy(claim_out, claim.v, 32);
    f_to_be32(&claim, be, F); k_update(&tr, be, 32);                    /* prover.rs:42 */
    for (unsigned t = 0; t <= degree; t++) {
        memcpy(round_polys_out + 4 * (size_t)t, acc[t].v, 32);
        f_to_be32(&acc[t], be, F); k_update(&tr, be, 32);
    }
    fe r = t_sample(&tr, F);
    memcpy(challenges_out, r.v, 32);
    for (int i = 0; i < nj; i++) jobs[i].r = r;
    gen_run(gen_fold_worker, jobs, nj);
    mt_rounds(F, &tr, (uint64_t *const *)tabs, m, n_vars, 1, degree, round_polys_out, challenges_out, n_threads);
    for (unsigned k = 0; k < m; k++) { if (finals_out) memcpy(finals_out + 4 * k, tabs[k], 32); free(tabs[k]); }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Sum of products  P(x) = sum_t prod_{k in term t} A_k(x)   (NOT in the reference,
 * whose ProductPoly is one product, product_poly.rs:4-10).  The loop is prover.rs:33-73 verbatim
 * with "prod_reduce().iter().sum()" read as the hypercube sum of P: append the sum, per round the
 * evaluations at t = 0..degree (fold of every table at t, evaluation_form.rs:68), a challenge, the
 * fold at it.  term_fac is n_terms rows of 8 table indices, term_len[t] of them used.
 * Written REFERENCE-SHAPED on purpose (a partial_evaluate clone per table and point, a materialised
 * table of P, a separate sum pass) so that it shares no structure with the GPU kernel it checks.
 * ---------------------------------------------------------------------------------------- */
static fe sop_table_sum(fe *const *tabs, size_t len, const uint8_t *term_len, const uint8_t *term_fac, unsigned n_terms,
                        const field_t *F) {
    fe *P = (fe *)malloc(len * sizeof(fe));
    for (size_t j = 0; j < len; j++) P[j] = f_zero();
    fe *prod = (fe *)malloc(len * sizeof(fe));
    for (unsigned t = 0; t < n_terms; t++) {
        memcpy(prod, tabs[term_fac[8 * t]], len * sizeof(fe));
        for (unsigned i = 1; i < term_len[t]; i++) {
            const fe *B = tabs[term_fac[8 * t + i]];
            for (size_t j = 0; j < len; j++) prod[j] = f_mul(&prod[j], &B[j], F);
        }
        for (size_t j = 0; j < len; j++) P[j] = f_add(&P[j], &prod[j], F);
    }
    fe acc = f_zero();
    for (size_t j = 0; j < len; j++) acc = f_add(&acc, &P[j], F);
    free(prod); free(P);
    return acc;
}
EXPORT int zko_sop_sum(int field, const uint64_t *const *tables, unsigned n_tables, unsigned n_vars,
                       const uint8_t *term_len, const uint8_t *term_fac, unsigned n_terms, uint64_t out[4]) {
    const field_t *F = &FIELDS[field]; (void)n_tables;
    fe s = sop_table_sum((fe *const *)tables, (size_t)1 << n_vars, term_len, term_fac, n_terms, F);
    memcpy(out, s.v, 32);
    return 0;
}
EXPORT int zko_prove_sop(int field, const uint64_t *const *tables, unsigned n_tables, unsigned n_vars,
                         const uint8_t *term_len, const uint8_t *term_fac, unsigned n_terms, unsigned degree,
                         const uint64_t sum[4], int absorb, uint64_t *round_polys_out, uint64_t *challenges_out,
                         uint64_t *finals_out) {
    const field_t *F = &FIELDS[field];
    keccak_t tr; k_init(&tr);
    size_t n = (size_t)1 << n_vars;
    fe **poly = (fe **)malloc(n_tables * sizeof(fe *));
    for (unsigned k = 0; k < n_tables; k++) { poly[k] = (fe *)malloc(n * sizeof(fe)); memcpy(poly[k], tables[k], n * sizeof(fe)); }
    uint8_t be[32];
    if (absorb) {
        for (unsigned k = 0; k < n_tables; k++) for (size_t j = 0; j < n; j++) { f_to_be32(&poly[k][j], be, F); k_update(&tr, be, 32); }
    }
    fe s; memcpy(s.v, sum, 32); f_to_be32(&s, be, F); k_update(&tr, be, 32);
    unsigned nv = n_vars;
    for (unsigned round = 0; round < n_vars; round++) {
        size_t half = (size_t)1 << (nv - 1);
        for (unsigned t = 0; t <= degree; t++) {
            fe ft = f_from_u64(t, F);
            fe **pe = (fe **)malloc(n_tables * sizeof(fe *));
            for (unsigned k = 0; k < n_tables; k++) pe[k] = mle_partial_evaluate(poly[k], nv, 0, &ft, 1, F);
            fe acc = sop_table_sum(pe, half, term_len, term_fac, n_terms, F);
            memcpy(round_polys_out + 4 * ((size_t)round * (degree + 1) + t), acc.v, 32);
            f_to_be32(&acc, be, F); k_update(&tr, be, 32);
            for (unsigned k = 0; k < n_tables; k++) free(pe[k]);
            free(pe);
        }
        fe r = t_sample(&tr, F);
        memcpy(challenges_out + 4 * (size_t)round, r.v, 32);
        for (unsigned k = 0; k < n_tables; k++) {
            fe *nx = mle_partial_evaluate(poly[k], nv, 0, &r, 1, F); free(poly[k]); poly[k] = nx;
        }
        nv--;
    }
    for (unsigned k = 0; k < n_tables; k++) { if (finals_out) memcpy(finals_out + 4 * k, poly[k][0].v, 32); free(poly[k]); }
    free(poly);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * SumcheckVerifier::verify_internal — sumcheck/src/verifier.rs:44-78, with
 * UnivariatePolynomial::interpolate (univariate_poly.rs:43-80, Lagrange over x = 0..D in
 * coefficient form) and Horner evaluate (:29-40).
 * Returns 0 = Ok, 3 = Err("verifier check failed: claimed_sum != p(0) + p(1)").
 * subclaim_sum_out / challenges_out receive SubClaim{sum, challenges}.
 * initial_poly_bytes != NULL => `verify` (absorbs them first, verifier.rs:21-22).
 * ---------------------------------------------------------------------------------------- */
static void poly_mul(const fe *a, int na, const fe *b, int nb, fe *out, const field_t *F) {
    for (int i = 0; i < na + nb - 1; i++) out[i] = f_zero();
    for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++) { fe t = f_mul(&a[i], &b[j], F); out[i + j] = f_add(&out[i + j], &t, F); }
}
static void interpolate(const fe *ys, int n, fe *coef, const field_t *F) {
    fe xs[17]; for (int i = 0; i < n; i++) xs[i] = f_from_u64((uint64_t)i, F);
    for (int i = 0; i < n; i++) coef[i] = f_zero();
    for (int li = 0; li < n; li++) {
        fe basis[17]; int nb = 1; basis[0] = f_one(F);
        for (int xi = 0; xi < n; xi++) {
            if (xi == li) continue;
            fe zero = f_zero();
            fe num[2]; num[0] = f_sub(&zero, &xs[xi], F); num[1] = f_one(F);
            fe dx = f_sub(&xs[li], &xs[xi], F); fe den = f_inv(&dx, F);
            fe scaled[2]; scaled[0] = f_mul(&num[0], &den, F); scaled[1] = f_mul(&num[1], &den, F);
            fe tmp[17]; poly_mul(basis, nb, scaled, 2, tmp, F); nb += 1; memcpy(basis, tmp, nb * sizeof(fe));
        }
        for (int i = 0; i < nb; i++) { fe t = f_mul(&basis[i], &ys[li], F); coef[i] = f_add(&coef[i], &t, F); }
    }
}
static fe horner(const fe *coef, int n, const fe *x, const field_t *F) {
    fe acc = f_zero();
    for (int i = n - 1; i >= 0; i--) { acc = f_mul(&acc, x, F); acc = f_add(&acc, &coef[i], F); }
    return acc;
}
EXPORT int zko_verify_internal(int field, const uint8_t *initial_poly_bytes, size_t n_initial_bytes,
                               const uint64_t sum[4], const uint64_t *round_polys, unsigned n_rounds, unsigned degree,
                               uint64_t subclaim_sum_out[4], uint64_t *challenges_out) {
    const field_t *F = &FIELDS[field];
    if (degree > 15) return -1;
    keccak_t tr; k_init(&tr);
    if (initial_poly_bytes) k_update(&tr, initial_poly_bytes, n_initial_bytes);
    uint8_t be[32]; fe claimed; memcpy(claimed.v, sum, 32);
    f_to_be32(&claimed, be, F); k_update(&tr, be, 32);                  /* :50 */
    for (unsigned r = 0; r < n_rounds; r++) {                           /* :54 */
        const fe *rp = (const fe *)(round_polys + 4 * (size_t)r * (degree + 1));
        for (unsigned t = 0; t <= degree; t++) { f_to_be32(&rp[t], be, F); k_update(&tr, be, 32); }   /* :56 */
        fe coef[17]; interpolate(rp, (int)degree + 1, coef, F);         /* :58 */
        fe zero = f_zero(), one = f_one(F);
        fe p0 = horner(coef, (int)degree + 1, &zero, F), p1 = horner(coef, (int)degree + 1, &one, F);
        fe s01 = f_add(&p0, &p1, F);
        if (!f_eq(&claimed, &s01)) return 3;                            /* :64-66 */
        fe ch = t_sample(&tr, F);                                       /* :69 */
        claimed = horner(coef, (int)degree + 1, &ch, F);                /* :70 */
        memcpy(challenges_out + 4 * (size_t)r, ch.v, 32);
    }
    memcpy(subclaim_sum_out, claimed.v, 32);
    return 0;
}

/* poly.to_bytes() of a ProductPoly (product_poly.rs:77-83): m*2^n*32 bytes, factor-major */
EXPORT void zko_to_bytes(int field, const uint64_t *evals, size_t count, uint8_t *out) {
    const field_t *F = &FIELDS[field];
    for (size_t j = 0; j < count; j++) f_to_be32(&((const fe *)evals)[j], out + 32 * j, F);
}

/* ------------------------------------------------------------------------------------------
 * fft / ifft — fft/src/lib.rs:4-61, REFERENCE-SHAPED (recursive, split_even_odd into fresh
 * vectors, omega.pow([i]) and omega.pow([i+n/2]) per butterfly, one field inversion per output
 * element in ifft).  log_n must be <= two_adicity (else returns 1: the reference's unwrap panic).
 * ---------------------------------------------------------------------------------------- */
static fe root_of_unity(const field_t *F, unsigned log_n) {   /* g^((p-1)/2^log_n) */
    uint64_t e[4]; memcpy(e, F->p, 32); e[0] -= 1;
    for (unsigned s = 0; s < log_n; s++) { for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63); e[3] >>= 1; }
    fe g = f_from_u64(F->gen, F);
    return f_pow(&g, e, 4, F);
}
static fe *fft_internal(fe *values, size_t n, fe omega, const field_t *F) {   /* :21-46; consumes values */
    if (n == 1) return values;
    fe *even = (fe *)malloc(n / 2 * sizeof(fe)), *odd = (fe *)malloc(n / 2 * sizeof(fe));
    for (size_t i = 0; i < n; i++) { if (i % 2 == 0) even[i / 2] = values[i]; else odd[i / 2] = values[i]; }   /* :48-61 */
    free(values);
    fe w2 = f_mul(&omega, &omega, F);
    fe *ee = fft_internal(even, n / 2, w2, F), *oe = fft_internal(odd, n / 2, w2, F);
    fe *out = (fe *)malloc(n * sizeof(fe));
    for (size_t i = 0; i < n / 2; i++) {                                /* :40-43 */
        uint64_t e1 = (uint64_t)i, e2 = (uint64_t)(i + n / 2);
        fe w1 = f_pow(&omega, &e1, 1, F), wb = f_pow(&omega, &e2, 1, F);
        fe t1 = f_mul(&w1, &oe[i], F), t2 = f_mul(&wb, &oe[i], F);
        out[i] = f_add(&ee[i], &t1, F); out[i + n / 2] = f_add(&ee[i], &t2, F);
    }
    free(ee); free(oe);
    return out;
}
EXPORT int zko_fft(int field, const uint64_t *in, uint64_t *out, unsigned log_n, int inverse) {
    const field_t *F = &FIELDS[field];
    if (log_n > F->two_adicity) return 1;
    size_t n = (size_t)1 << log_n;
    fe omega = root_of_unity(F, log_n);                                 /* :6 */
    if (inverse) omega = f_inv(&omega, F);                              /* :14 */
    fe *v = (fe *)malloc(n * sizeof(fe)); memcpy(v, in, n * sizeof(fe));
    fe *r = fft_internal(v, n, omega, F);
    if (inverse) {                                                      /* :15-18: inverse() per element */
        fe fn = f_from_u64((uint64_t)n, F);
        for (size_t i = 0; i < n; i++) { fe ninv = f_inv(&fn, F); r[i] = f_mul(&r[i], &ninv, F); }
    }
    memcpy(out, r, n * sizeof(fe)); free(r);
    return 0;
}
/* Streamlined iterative radix-2 (bit-reverse + DIT, incremental twiddles): same outputs, used to
 * cross-check the GPU at sizes where the reference-shaped recursion is too slow. */
EXPORT int zko_fft_fast(int field, const uint64_t *in, uint64_t *out, unsigned log_n, int inverse) {
    const field_t *F = &FIELDS[field];
    if (log_n > F->two_adicity) return 1;
    size_t n = (size_t)1 << log_n;
    fe *a = (fe *)out;
    for (size_t i = 0; i < n; i++) {
        size_t r = 0; for (unsigned b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
        memcpy(&a[r], in + 4 * i, 32);
    }
    fe omega = root_of_unity(F, log_n); if (inverse) omega = f_inv(&omega, F);
    fe *tw = (fe *)malloc((n / 2 ? n / 2 : 1) * sizeof(fe));
    tw[0] = f_one(F); for (size_t i = 1; i < n / 2; i++) tw[i] = f_mul(&tw[i - 1], &omega, F);
    for (unsigned s = 1; s <= log_n; s++) {
        size_t len = (size_t)1 << s, half = len / 2, step = n / len;
        for (size_t base = 0; base < n; base += len)
            for (size_t j = 0; j < half; j++) {
                fe t = f_mul(&tw[j * step], &a[base + j + half], F);
                fe u = a[base + j];
                a[base + j] = f_add(&u, &t, F); a[base + j + half] = f_sub(&u, &t, F);
            }
    }
    free(tw);
    if (inverse) { fe fn = f_from_u64((uint64_t)n, F); fe ninv = f_inv(&fn, F); for (size_t i = 0; i < n; i++) a[i] = f_mul(&a[i], &ninv, F); }
    return 0;
}
