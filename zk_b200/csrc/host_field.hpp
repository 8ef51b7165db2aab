// host_field.hpp — host-side prime-field arithmetic for the parts of the path that stay on the CPU:
// Fiat-Shamir challenge derivation (`F::from_be_bytes_mod_order`, transcript/src/lib.rs:27-30),
// transcript serialisation (`into_bigint().to_bytes_be()`, sumcheck/src/lib.rs:23-29) and the
// verifier's O(n*D^2) interpolation (sumcheck/src/verifier.rs:44-78, polynomial/src/univariate_poly.rs:29-80).
// Same memory format as the device: 4 LE u64 limbs, Montgomery form, fully reduced.
#pragma once
#include <cstdint>
#include <cstring>

namespace zk {
namespace host {

typedef unsigned __int128 u128;

struct FieldParams {
    uint64_t p[4];
    uint64_t one[4];  // R mod p
    uint64_t r2[4];   // R^2 mod p
    uint64_t inv;     // -p^-1 mod 2^64
    uint64_t generator;
    unsigned two_adicity;
};

inline const FieldParams& params(int field) {
    static const FieldParams k[2] = {
        {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
         {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL},
         {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL},
         0xfffffffeffffffffULL, 7, 32},
        {{0x0a11800000000001ULL, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL, 0x12ab655e9a2ca556ULL},
         {0x7d1c7ffffffffff3ULL, 0x7257f50f6ffffff2ULL, 0x16d81575512c0feeULL, 0x0d4bda322bbb9a9dULL},
         {0x25d577bab861857bULL, 0xcc2c27b58860591fULL, 0xa7cc008fe5dc8593ULL, 0x011fdae7eff1c939ULL},
         0x0a117fffffffffffULL, 22, 47},
    };
    return k[field];
}

struct El {
    uint64_t v[4];
    bool operator==(const El& o) const { return std::memcmp(v, o.v, 32) == 0; }
    bool operator!=(const El& o) const { return !(*this == o); }
};

class Field {
   public:
    explicit Field(int id) : id_(id), P(params(id)) {}
    int id() const { return id_; }
    unsigned two_adicity() const { return P.two_adicity; }

    El zero() const { return El{{0, 0, 0, 0}}; }
    El one() const { El r; std::memcpy(r.v, P.one, 32); return r; }

    bool geq_p(const uint64_t a[4]) const {
        for (int i = 3; i >= 0; i--) {
            if (a[i] != P.p[i]) return a[i] > P.p[i];
        }
        return true;
    }
    void sub_p(uint64_t a[4]) const {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)a[i] - P.p[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    El add(const El& a, const El& b) const {
        El r; u128 c = 0;
        for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    El sub(const El& a, const El& b) const {
        El r; uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)a.v[i] - b.v[i] - borrow;
            r.v[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + P.p[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    El neg(const El& a) const { return sub(zero(), a); }
    // Montgomery product: coarsely integrated operand scanning on 4 x 64-bit limbs, branch free (it sits on the round loop's
    // latency chain: challenge multiples, round-polynomial evaluation, transcript serialisation — a few dozen per round)
    El mul(const El& a, const El& b) const {
        uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        for (int i = 0; i < 4; i++) {
            const uint64_t bi = b.v[i];
            u128 c = (u128)a.v[0] * bi + t0;
            const uint64_t lo0 = (uint64_t)c;
            c = (u128)a.v[1] * bi + t1 + (uint64_t)(c >> 64);
            const uint64_t lo1 = (uint64_t)c;
            c = (u128)a.v[2] * bi + t2 + (uint64_t)(c >> 64);
            const uint64_t lo2 = (uint64_t)c;
            c = (u128)a.v[3] * bi + t3 + (uint64_t)(c >> 64);
            const uint64_t lo3 = (uint64_t)c;
            const u128 top = (u128)t4 + (uint64_t)(c >> 64);  // < 2^65
            const uint64_t m = lo0 * P.inv;
            c = (u128)m * P.p[0] + lo0;                        // low word becomes 0
            c = (u128)m * P.p[1] + lo1 + (uint64_t)(c >> 64);
            t0 = (uint64_t)c;
            c = (u128)m * P.p[2] + lo2 + (uint64_t)(c >> 64);
            t1 = (uint64_t)c;
            c = (u128)m * P.p[3] + lo3 + (uint64_t)(c >> 64);
            t2 = (uint64_t)c;
            const u128 hi = top + (uint64_t)(c >> 64);
            t3 = (uint64_t)hi;
            t4 = (uint64_t)(hi >> 64);
        }
        El r{{t0, t1, t2, t3}};
        if (t4 || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    El from_canonical(const uint64_t c[4]) const {  // c < p
        El a, r2; std::memcpy(a.v, c, 32); std::memcpy(r2.v, P.r2, 32);
        return mul(a, r2);
    }
    El from_u64(uint64_t x) const { uint64_t c[4] = {x, 0, 0, 0}; return from_canonical(c); }
    void to_canonical(const El& a, uint64_t out[4]) const {
        El raw_one{{1, 0, 0, 0}};
        El r = mul(a, raw_one);
        std::memcpy(out, r.v, 32);
    }
    El pow(const El& a, const uint64_t* e, int limbs) const {
        El r = one();
        for (int i = limbs - 1; i >= 0; i--)
            for (int b = 63; b >= 0; b--) { r = mul(r, r); if ((e[i] >> b) & 1) r = mul(r, a); }
        return r;
    }
    El inverse(const El& a) const {  // a != 0: a^(p-2)
        uint64_t e[4]; std::memcpy(e, P.p, 32); e[0] -= 2;
        return pow(a, e, 4);
    }
    // g^((p-1)/2^log_n): ark-ff FftField::get_root_of_unity(2^log_n); caller checks log_n <= two_adicity
    El root_of_unity(unsigned log_n) const {
        uint64_t e[4]; std::memcpy(e, P.p, 32); e[0] -= 1;
        for (unsigned s = 0; s < log_n; s++) {
            for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
            e[3] >>= 1;
        }
        return pow(from_u64(P.generator), e, 4);
    }
    void to_be32(const El& a, uint8_t out[32]) const {
        uint64_t c[4]; to_canonical(a, c);
        for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) out[31 - (8 * i + b)] = (uint8_t)(c[i] >> (8 * b));
    }
    El from_be32_mod_order(const uint8_t in[32]) const {
        uint64_t c[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) c[i] |= (uint64_t)in[31 - (8 * i + b)] << (8 * b);
        while (geq_p(c)) sub_p(c);  // 2^256 < 3p (381), < 14p (377)
        return from_canonical(c);
    }
    bool is_canonical(const El& a) const { return !geq_p(a.v); }

   private:
    int id_;
    const FieldParams& P;
};

// Multiples of a launch-wide multiplier for the device's fe_mul_fixed: out[i] = r * 2^(32 i + 64) mod p as plain
// (canonical, non-Montgomery) 8x32-bit limbs.  `r_mont` is r in Montgomery form.
inline void fixed_mul_table(const Field& F, const El& r_mont, uint32_t out[8][8]) {
    El two32 = F.from_u64((uint64_t)1 << 32);
    El cur = F.mul(r_mont, F.mul(two32, two32));  // r * 2^64 (Montgomery form)
    for (int i = 0; i < 8; i++) {
        uint64_t c[4];
        F.to_canonical(cur, c);
        for (int j = 0; j < 4; j++) {
            out[i][2 * j] = (uint32_t)c[j];
            out[i][2 * j + 1] = (uint32_t)(c[j] >> 32);
        }
        cur = F.mul(cur, two32);
    }
}

// Multiples of a launch-wide multiplier for the device's fe_mul_fixed_f64 (field_f64.cuh):
// out[i][j] = limb j (32 bits, as a double) of r * 2^(16 i + 32) mod p, canonical.  `r_mont` is r in Montgomery form.
inline void fixed_mul_table_f64(const Field& F, const El& r_mont, double out[16][8]) {
    const El two16 = F.from_u64((uint64_t)1 << 16);
    El cur = F.mul(r_mont, F.from_u64((uint64_t)1 << 32));
    for (int i = 0; i < 16; i++) {
        uint64_t c[4];
        F.to_canonical(cur, c);
        for (int j = 0; j < 4; j++) {
            out[i][2 * j] = (double)(uint32_t)c[j];
            out[i][2 * j + 1] = (double)(uint32_t)(c[j] >> 32);
        }
        cur = F.mul(cur, two16);
    }
}

}  // namespace host
}  // namespace zk
