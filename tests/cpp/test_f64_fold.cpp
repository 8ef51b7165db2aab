// Host check of the FP64 fixed-multiplier product (zk_b200/csrc/field_f64.cuh): the column sums are plain IEEE-754
// binary64 arithmetic on integers < 2^53, so the device function's arithmetic can be replayed bit for bit on the CPU.
// Compares  (sum_j col_j 2^(32j) + m p) / 2^32, conditionally reduced,  with the word-serial host Montgomery product
// (host_field.hpp) for random and extreme operands, both fields.
// Build: g++ -std=c++17 -O2 -ffp-contract=off -I zk_b200/csrc tests/cpp/test_f64_fold.cpp -o build/test_f64_fold
#include <cstdio>
#include <cstdlib>

#define __align__(n) __attribute__((aligned(n)))
#define __host__
#define __device__
#include "field_f64.cuh"
#include "host_field.hpp"

using namespace zk;
typedef unsigned __int128 u128;

static uint64_t smix(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
    return z;
}

// the integer tail of fe_mul_fixed_f64, restated with 64-bit words
static host::El tail(const host::Field& F, const host::FieldParams& P, const double* col, double* max_col, uint32_t top_addend = 0) {
    uint32_t V[10] = {0};
    V[8] = top_addend;
    uint64_t carry = 0;
    for (int j = 0; j < 8; j++) {
        uint64_t bits;
        std::memcpy(&bits, &col[j], 8);
        if ((bits >> 52) != 0x433) { std::printf("column %d left the 2^52 binade\n", j); std::exit(1); }
        const uint64_t v = bits & ((1ULL << 52) - 1);
        if ((double)v > *max_col) *max_col = (double)v;
        // V += v << 32j
        u128 acc = (u128)v + V[j] + ((u128)V[j + 1] << 32);
        V[j] = (uint32_t)acc;
        V[j + 1] = (uint32_t)(acc >> 32);
        carry = (uint64_t)(acc >> 64);
        for (int k = j + 2; carry && k < 10; k++) { uint64_t s = (uint64_t)V[k] + carry; V[k] = (uint32_t)s; carry = s >> 32; }
    }
    // one Montgomery row with 32-bit words: m = -V0 (p == 1 mod 2^32)
    const uint32_t m = 0u - V[0];
    uint64_t c = 0;
    for (int k = 0; k < 8; k++) {
        const uint32_t pk = (uint32_t)(P.p[k / 2] >> (32 * (k & 1)));
        c += (uint64_t)m * pk + V[k];
        V[k] = (uint32_t)c;
        c >>= 32;
    }
    for (int k = 8; k < 10; k++) { c += V[k]; V[k] = (uint32_t)c; c >>= 32; }
    if (V[0] != 0 || V[9] != 0) { std::printf("row did not clear / overflow\n"); std::exit(1); }
    host::El r;
    for (int k = 0; k < 4; k++) r.v[k] = (uint64_t)V[1 + 2 * k] | ((uint64_t)V[2 + 2 * k] << 32);
    if (F.geq_p(r.v)) F.sub_p(r.v);
    if (top_addend == 0 && F.geq_p(r.v)) { std::printf("result >= 2p\n"); std::exit(1); }
    if (F.geq_p(r.v)) F.sub_p(r.v);
    if (F.geq_p(r.v)) { std::printf("result >= 3p\n"); std::exit(1); }
    return r;
}

int main() {
    long bad = 0, total = 0;
    for (int fid = 0; fid < 2; fid++) {
        host::Field F(fid);
        const host::FieldParams& P = host::params(fid);
        double max_col = 0;
        for (int rr = 0; rr < 6; rr++) {
            host::El r;
            if (rr == 0) r = F.zero();
            else if (rr == 1) r = F.one();
            else if (rr == 2) r = F.neg(F.one());  // p - 1
            else { uint64_t c[4]; for (int l = 0; l < 4; l++) c[l] = smix(77 * rr + l); c[3] &= 0x0FFFFFFFFFFFFFFFULL; r = F.from_canonical(c); }
            FixedMulF64 tab;
            host::fixed_mul_table_f64(F, r, tab.t);
            for (int i = 0; i < 200000; i++) {
                host::El x;
                for (int l = 0; l < 4; l++) x.v[l] = smix(0x1234 + (uint64_t)rr * 1000003 + (uint64_t)i * 4 + l);
                x.v[3] &= 0x0FFFFFFFFFFFFFFFULL;
                if (i == 0) x = F.zero();
                if (i == 1) { std::memcpy(x.v, P.p, 32); x.v[0] -= 1; }
                if (i == 2) { for (int l = 0; l < 4; l++) x.v[l] = ~0ULL; }  // 2^256 - 1: every half at its maximum
                if (i == 3) { x = F.zero(); x.v[0] = 0xffffffffULL; }
                uint32_t xv[8];
                for (int l = 0; l < 4; l++) { xv[2 * l] = (uint32_t)x.v[l]; xv[2 * l + 1] = (uint32_t)(x.v[l] >> 32); }
                double col[8];
                f64_columns(col, xv, tab);
                host::El got = tail(F, P, col, &max_col);
                // expected: x * r as Montgomery product of x with rR (for i == 2 x is not reduced: reduce it first)
                host::El xr = x;
                while (F.geq_p(xr.v)) F.sub_p(xr.v);
                host::El want = F.mul(xr, r);
                total++;
                if (got != want) { if (bad < 5) std::printf("mismatch field %d r %d i %d\n", fid, rr, i); bad++; }
                // the fused fold l + r (h - l) with l riding in the columns' start values (fe_fold_fixed_f64_x2)
                host::El l = xr, h;
                for (int k = 0; k < 4; k++) h.v[k] = smix(0x777 + (uint64_t)i * 4 + k);
                h.v[3] &= 0x0FFFFFFFFFFFFFFFULL;
                if (i == 1) h = F.zero();  // l = p-1, h = 0
                if (i == 4) { l = F.zero(); std::memcpy(h.v, P.p, 32); h.v[0] -= 1; }
                if (i == 5) { std::memcpy(l.v, P.p, 32); l.v[0] -= 1; h = l; }
                host::El d = F.sub(h, l);
                uint32_t dv[8], lv[8];
                for (int k = 0; k < 4; k++) { dv[2 * k] = (uint32_t)d.v[k]; dv[2 * k + 1] = (uint32_t)(d.v[k] >> 32); lv[2 * k] = (uint32_t)l.v[k]; lv[2 * k + 1] = (uint32_t)(l.v[k] >> 32); }
                double colf[1][8];
                const uint32_t* xs[1] = {dv};
                const uint32_t* as[1] = {lv};
                f64_columns_n<1>(colf, xs, tab, as);
                double dummy = 0;
                host::El gotf = tail(F, P, colf[0], &dummy, lv[7]);
                host::El wantf = F.sub(l, F.mul(F.sub(l, h), r));  // evaluation_form.rs:68
                total++;
                if (gotf != wantf) { if (bad < 5) std::printf("fold mismatch field %d r %d i %d\n", fid, rr, i); bad++; }
            }
        }
        std::printf("field %d: max column sum 2^%.3f (must stay below 2^52)\n", fid, __builtin_log2(max_col));
        if (max_col >= 4503599627370496.0) bad++;
    }
    std::printf("f64 fixed-multiplier product: %ld checked, %ld mismatches\n", total, bad);
    return bad != 0;
}
