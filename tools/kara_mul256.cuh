// kara.cuh — 256 x 256 -> 512-bit product with one level of (subtractive) Karatsuba: three 128 x 128 products
// (48 wide multiplies) instead of the 64 of the schoolbook rows in field.cuh, the difference paid in carry adds that
// run on the ALU pipe while the multiplier pipe is the binding one.
// Written on single-instruction carry primitives (the CGBN style) so that the very same source is replayed on the
// host (tools/kara_host_test.cpp defines ZK_KARA_HOST and models the carry flag) — the index bookkeeping is checked
// on the CPU against a plain big-integer product before a GPU ever sees it.
#pragma once
#include <cstdint>

namespace zk {
namespace kara {

#ifdef ZK_KARA_HOST
static thread_local uint32_t cf = 0;  // the PTX carry flag CC.CF
#define ZK_KARA_FN static inline
ZK_KARA_FN void add_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; r = (uint32_t)t; cf = (uint32_t)(t >> 32); }
ZK_KARA_FN void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf; r = (uint32_t)t; cf = (uint32_t)(t >> 32); }
ZK_KARA_FN void addc(uint32_t& r, uint32_t a, uint32_t b) { r = a + b + cf; }
ZK_KARA_FN void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; r = (uint32_t)t; cf = (uint32_t)(t >> 63); }
ZK_KARA_FN void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf; r = (uint32_t)t; cf = (uint32_t)(t >> 63); }
ZK_KARA_FN void subc(uint32_t& r, uint32_t a, uint32_t b) { r = a - b - cf; }
ZK_KARA_FN void mul_lo(uint32_t& r, uint32_t a, uint32_t b) { r = a * b; }
ZK_KARA_FN void mul_hi(uint32_t& r, uint32_t a, uint32_t b) { r = (uint32_t)(((uint64_t)a * b) >> 32); }
ZK_KARA_FN void mad_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; r = (uint32_t)t; cf = (uint32_t)(t >> 32); }
ZK_KARA_FN void madc_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + cf; r = (uint32_t)t; cf = (uint32_t)(t >> 32); }
ZK_KARA_FN void madc_hi_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + cf; r = (uint32_t)t; cf = (uint32_t)(t >> 32); }
ZK_KARA_FN void madc_hi(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { r = (uint32_t)(((uint64_t)a * b) >> 32) + c + cf; }
#else
#define ZK_KARA_FN __device__ __forceinline__
// PTX: sub.cc writes the borrow to CC.CF and subc computes a - (b + CF), as the host model above does.
ZK_KARA_FN void add_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void addc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void subc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void mul_lo(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void mul_hi(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
ZK_KARA_FN void mad_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
ZK_KARA_FN void madc_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
ZK_KARA_FN void madc_hi_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
ZK_KARA_FN void madc_hi(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
#endif

// out[0..7] = a[0..3] * b[0..3]: the even / odd carry-chain rows of fe_mul_wide at half width (16 wide multiplies).
// X holds the product pairs of the even limbs of a (columns c, c+1, c+2, c+3), Y those of the odd limbs one column up;
// after every row the lowest column is final and the roles swap.
ZK_KARA_FN void mul128(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    uint32_t X[4], Y[4];
    mul_lo(Y[0], a[1], b[0]); mul_hi(Y[1], a[1], b[0]); mul_lo(Y[2], a[3], b[0]); mul_hi(Y[3], a[3], b[0]);
    mul_lo(X[0], a[0], b[0]); mul_hi(X[1], a[0], b[0]); mul_lo(X[2], a[2], b[0]); mul_hi(X[3], a[2], b[0]);
    out[0] = X[0];
    uint32_t* E = X;  // array whose [0] was just emitted (its [1] still has to join the other array's [0])
    uint32_t* O = Y;
#pragma unroll
    for (int i = 1; i < 4; i++) {
        // E shifts two columns up and takes the odd products of this row; its column-(c+1) limb joins O[0] first
        add_cc(O[0], O[0], E[1]);
        madc_lo_cc(E[0], a[1], b[i], E[2]);
        madc_hi_cc(E[1], a[1], b[i], E[3]);
        madc_lo_cc(E[2], a[3], b[i], 0);
        madc_hi(E[3], a[3], b[i], 0);
        // O takes the even products of this row; the carry out goes to E's top column
        mad_lo_cc(O[0], a[0], b[i], O[0]);
        madc_hi_cc(O[1], a[0], b[i], O[1]);
        madc_lo_cc(O[2], a[2], b[i], O[2]);
        madc_hi_cc(O[3], a[2], b[i], O[3]);
        addc(E[3], E[3], 0);
        out[i] = O[0];
        uint32_t* t = E; E = O; O = t;
    }
    // E[0] was emitted as column 3; E[1..3] are columns 4..6, O[0..3] columns 4..7
    add_cc(out[4], O[0], E[1]);
    addc_cc(out[5], O[1], E[2]);
    addc_cc(out[6], O[2], E[3]);
    addc(out[7], O[3], 0);
}

// d = |x - y| over four limbs; returns the sign mask (all ones when x < y)
ZK_KARA_FN uint32_t absdiff128(uint32_t* d, const uint32_t* x, const uint32_t* y) {
    uint32_t s;
    sub_cc(d[0], x[0], y[0]); subc_cc(d[1], x[1], y[1]); subc_cc(d[2], x[2], y[2]); subc_cc(d[3], x[3], y[3]);
    subc(s, 0, 0);  // 0 - 0 - borrow
    // two's complement when negative: (d ^ s) + (s & 1)
    add_cc(d[0], d[0] ^ s, s & 1u); addc_cc(d[1], d[1] ^ s, 0); addc_cc(d[2], d[2] ^ s, 0); addc(d[3], d[3] ^ s, 0);
    return s;
}

// out[0..15] = a[0..7] * b[0..7]
ZK_KARA_FN void mul256(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    uint32_t z0[8], z2[8], m[8], da[4], db[4], t[9];
    mul128(z0, a, b);
    mul128(z2, a + 4, b + 4);
    const uint32_t sa = absdiff128(da, a, a + 4), sb = absdiff128(db, b, b + 4);
    mul128(m, da, db);
    // middle term z1 = z0 + z2 - (a0 - a1)(b0 - b1) = z0 + z2 -/+ m: subtract m when the signs agree
    const uint32_t neg = ~(sa ^ sb);
    add_cc(t[0], z0[0], z2[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) addc_cc(t[i], z0[i], z2[i]);
    addc(t[8], 0, 0);
    uint32_t dummy;
    add_cc(dummy, neg, 1u);  // carry in = 1 when subtracting (two's complement of m over nine limbs)
#pragma unroll
    for (int i = 0; i < 8; i++) addc_cc(t[i], t[i], m[i] ^ neg);
    addc(t[8], t[8], neg);
    // out = z0 + t * 2^128 + z2 * 2^256
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = z0[i];
    add_cc(out[4], z0[4], t[0]);
    addc_cc(out[5], z0[5], t[1]);
    addc_cc(out[6], z0[6], t[2]);
    addc_cc(out[7], z0[7], t[3]);
    addc_cc(out[8], z2[0], t[4]);
    addc_cc(out[9], z2[1], t[5]);
    addc_cc(out[10], z2[2], t[6]);
    addc_cc(out[11], z2[3], t[7]);
    addc_cc(out[12], z2[4], t[8]);
    addc_cc(out[13], z2[5], 0);
    addc_cc(out[14], z2[6], 0);
    addc(out[15], z2[7], 0);
}

}  // namespace kara
}  // namespace zk
