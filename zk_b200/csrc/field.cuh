// field.cuh — 256-bit prime-field arithmetic for sm_100a (8 x 32-bit limbs, Montgomery form, R = 2^256).
//
// Replaces, on the device, the third-party arithmetic the reference calls on every hot-loop
// iteration: ark_ff::Fp<MontBackend<FrConfig,4>,4>::{mul,add,sub} (call sites:
// polynomial/src/multilinear/evaluation_form.rs:61-68, polynomial/src/product_poly.rs:70,
// sumcheck/src/prover.rs:51-54, fft/src/lib.rs:41-42).  Values in memory are bit-identical to
// ark-ff's: 4 little-endian u64 limbs (= 8 little-endian u32 limbs), Montgomery form, fully reduced.
//
// The multiplier is an operand-scanning (CIOS) Montgomery product on two interleaved accumulator
// rows ("even" columns / "odd" columns) so that every 32x32->64 partial product lands on an aligned
// register pair and the whole row is ONE carry chain: ptxas lowers each mad.lo.cc/madc.hi.cc pair to
// a single IMAD.WIDE.U32(.X) with predicate carries.  Both supported moduli are == 1 (mod 2^32), so
// the per-row Montgomery factor is m = -t0 (no multiply) and m*p_0 needs no multiply either.
#pragma once
#include <cstdint>

namespace zk {

struct __align__(32) Fe {
    uint32_t v[8];
};

// ---- field descriptions -------------------------------------------------------------------
// BLS12-381 scalar field (ark-bls12-381 0.5.0 Fr): 255 bits.
struct Fr381 {
    static constexpr int ID = 0;
    static constexpr unsigned TWO_ADICITY = 32;
    static constexpr uint64_t GENERATOR = 7;
    __host__ __device__ static constexpr uint32_t p(int i) {
        constexpr uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return t[i];
    }
    __host__ __device__ static constexpr uint32_t one(int i) {  // R mod p
        constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return t[i];
    }
    __host__ __device__ static constexpr uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return t[i];
    }
    __host__ __device__ static constexpr uint32_t k288(int i) {  // 2^288 mod p (plain integer)
        constexpr uint32_t t[8] = {0xcaaf6b13u, 0x355094eau, 0x69a568efu, 0xf6b10cb3u,
                                   0x40cc3869u, 0xe2c926a6u, 0xed269aadu, 0x736a6d3bu};
        return t[i];
    }
};
// BLS12-377 scalar field (ark-bls12-377 0.5.0 Fr): 253 bits — the field of the reference's fft test.
struct Fr377 {
    static constexpr int ID = 1;
    static constexpr unsigned TWO_ADICITY = 47;
    static constexpr uint64_t GENERATOR = 22;
    __host__ __device__ static constexpr uint32_t p(int i) {
        constexpr uint32_t t[8] = {0x00000001u, 0x0a118000u, 0xd0000001u, 0x59aa76feu,
                                   0x5c37b001u, 0x60b44d1eu, 0x9a2ca556u, 0x12ab655eu};
        return t[i];
    }
    __host__ __device__ static constexpr uint32_t one(int i) {
        constexpr uint32_t t[8] = {0xfffffff3u, 0x7d1c7fffu, 0x6ffffff2u, 0x7257f50fu,
                                   0x512c0feeu, 0x16d81575u, 0x2bbb9a9du, 0x0d4bda32u};
        return t[i];
    }
    __host__ __device__ static constexpr uint32_t r2(int i) {
        constexpr uint32_t t[8] = {0xb861857bu, 0x25d577bau, 0x8860591fu, 0xcc2c27b5u,
                                   0xe5dc8593u, 0xa7cc008fu, 0xeff1c939u, 0x011fdae7u};
        return t[i];
    }
    __host__ __device__ static constexpr uint32_t k288(int i) {
        constexpr uint32_t t[8] = {0x49adb84fu, 0x2f667ff2u, 0xef9e8ae2u, 0xe4a46e13u,
                                   0xe7d8fb0eu, 0x29d63165u, 0xbdb3a756u, 0x00342799u};
        return t[i];
    }
};

// Precomputed multiples of a launch-wide multiplier (see fe_mul_fixed): v[i] = r * 2^(32 i + 64) mod p.
struct FixedMul {
    uint32_t v[8][8];
};

#ifdef __CUDACC__

// ---- 256-bit global memory access (sm_100+: LDG.E.ENL2.256 / STG.E.ENL2.256) ------------------
// One instruction moves a whole 32-byte element, so the reference's AoS layout is also the
// perfectly coalesced layout: a warp touches 1024 contiguous bytes per instruction.
__device__ __forceinline__ Fe ld_fe(const Fe* p) {
    Fe r;
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                   "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}
// streaming variant for data that is read exactly once (do not keep in L1)
__device__ __forceinline__ Fe ld_fe_stream(const Fe* p) {
    Fe r;
    asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                   "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_fe(Fe* p, const Fe& r) {
    asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
                 "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p)
                 : "memory");
}

template <class F>
__device__ __forceinline__ Fe fe_zero() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
template <class F>
__device__ __forceinline__ Fe fe_one() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = F::one(i);
    return r;
}

// r = a + b (mod p), a,b in [0,p) -> [0,p)
template <class F>
__device__ __forceinline__ Fe fe_add(const Fe& a, const Fe& b) {
    Fe s, t;
    asm("add.cc.u32 %0,%8,%16;\n\taddc.cc.u32 %1,%9,%17;\n\taddc.cc.u32 %2,%10,%18;\n\taddc.cc.u32 %3,%11,%19;\n\t"
        "addc.cc.u32 %4,%12,%20;\n\taddc.cc.u32 %5,%13,%21;\n\taddc.cc.u32 %6,%14,%22;\n\taddc.u32 %7,%15,%23;"
        : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    uint32_t borrow;
    asm("sub.cc.u32 %0,%9,%17;\n\tsubc.cc.u32 %1,%10,%18;\n\tsubc.cc.u32 %2,%11,%19;\n\tsubc.cc.u32 %3,%12,%20;\n\t"
        "subc.cc.u32 %4,%13,%21;\n\tsubc.cc.u32 %5,%14,%22;\n\tsubc.cc.u32 %6,%15,%23;\n\tsubc.cc.u32 %7,%16,%24;\n\t"
        "subc.u32 %8,0,0;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]),
          "=r"(borrow)
        : "r"(s.v[0]), "r"(s.v[1]), "r"(s.v[2]), "r"(s.v[3]), "r"(s.v[4]), "r"(s.v[5]), "r"(s.v[6]), "r"(s.v[7]),
          "r"(F::p(0)), "r"(F::p(1)), "r"(F::p(2)), "r"(F::p(3)), "r"(F::p(4)), "r"(F::p(5)), "r"(F::p(6)), "r"(F::p(7)));
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? s.v[i] : t.v[i];
    return r;
}

// r = a - b (mod p), a,b in [0,p) -> [0,p)
template <class F>
__device__ __forceinline__ Fe fe_sub(const Fe& a, const Fe& b) {
    Fe d;
    uint32_t borrow;
    asm("sub.cc.u32 %0,%9,%17;\n\tsubc.cc.u32 %1,%10,%18;\n\tsubc.cc.u32 %2,%11,%19;\n\tsubc.cc.u32 %3,%12,%20;\n\t"
        "subc.cc.u32 %4,%13,%21;\n\tsubc.cc.u32 %5,%14,%22;\n\tsubc.cc.u32 %6,%15,%23;\n\tsubc.cc.u32 %7,%16,%24;\n\t"
        "subc.u32 %8,0,0;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7]),
          "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // borrow is 0 or 0xffffffff: add back p & borrow
    Fe r;
    asm("add.cc.u32 %0,%8,%16;\n\taddc.cc.u32 %1,%9,%17;\n\taddc.cc.u32 %2,%10,%18;\n\taddc.cc.u32 %3,%11,%19;\n\t"
        "addc.cc.u32 %4,%12,%20;\n\taddc.cc.u32 %5,%13,%21;\n\taddc.cc.u32 %6,%14,%22;\n\taddc.u32 %7,%15,%23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]), "r"(d.v[7]),
          "r"(F::p(0) & borrow), "r"(F::p(1) & borrow), "r"(F::p(2) & borrow), "r"(F::p(3) & borrow),
          "r"(F::p(4) & borrow), "r"(F::p(5) & borrow), "r"(F::p(6) & borrow), "r"(F::p(7) & borrow));
    return r;
}

// x in [0, 2p) (and < 2^256) -> [0,p)
template <class F>
__device__ __forceinline__ Fe fe_reduce_once(const Fe& x) {
    Fe t;
    uint32_t borrow;
    asm("sub.cc.u32 %0,%9,%17;\n\tsubc.cc.u32 %1,%10,%18;\n\tsubc.cc.u32 %2,%11,%19;\n\tsubc.cc.u32 %3,%12,%20;\n\t"
        "subc.cc.u32 %4,%13,%21;\n\tsubc.cc.u32 %5,%14,%22;\n\tsubc.cc.u32 %6,%15,%23;\n\tsubc.cc.u32 %7,%16,%24;\n\t"
        "subc.u32 %8,0,0;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]),
          "=r"(borrow)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "r"(F::p(0)), "r"(F::p(1)), "r"(F::p(2)), "r"(F::p(3)), "r"(F::p(4)), "r"(F::p(5)), "r"(F::p(6)), "r"(F::p(7)));
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? x.v[i] : t.v[i];
    return r;
}

// ---- Montgomery multiplication ---------------------------------------------------------------
namespace detail {
// X[0..7] += {a0,a2,a4,a6} * b as one carry chain; the carry out is added to `top`.
__device__ __forceinline__ void cmad4(uint32_t* X, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b,
                                      uint32_t& top) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(top)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}
// same, carry out provably zero (dropped)
__device__ __forceinline__ void cmad4_nocarry(uint32_t* X, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6,
                                              uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}
// X = {a0,a2,a4,a6} * b (first row)
__device__ __forceinline__ void mul4(uint32_t* X, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(X[0]), "=r"(X[1]), "=r"(X[2]), "=r"(X[3]), "=r"(X[4]), "=r"(X[5]), "=r"(X[6]), "=r"(X[7])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}
// One-column right shift fused into the multiply-accumulate: the limb that falls off the bottom of
// the other row (Y[1], same column as x0) is added into x0 first and its carry enters the chain;
// then Y[j] = {a1,a3,a5,a7}*b + Y[j+2].
__device__ __forceinline__ void madc4_rshift(uint32_t* Y, uint32_t& x0, uint32_t a1, uint32_t a3, uint32_t a5,
                                             uint32_t a7, uint32_t b) {
    asm("add.cc.u32 %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %2;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
        "madc.hi.u32 %7, %12, %13, 0;"
        : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7]), "+r"(x0)
        : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}
// Reduction half-rows.  Both moduli have p_0 = 1, so X[0] + m*p_0 is just X[0] + m = 0 with carry
// (X[0] != 0): two carry adds instead of a multiply.  For BLS12-381 Fr also p_1 = 2^32 - 1, and
// m*(2^32-1) = ((m - c) << 32) + t0 with t0 = -m, c = (t0 != 0): two more adds instead of a multiply.
// IMAD.WIDE.U32 issues at half the IMAD rate on sm_100a (32 per clock per SM, measured) and is the
// binding pipe of every kernel here while the ALU pipe idles, so: 112 (381) / 120 (377) wide multiplies
// per field multiplication instead of 128.
__device__ __forceinline__ void redc_even(uint32_t* X, uint32_t m, uint32_t p2, uint32_t p4, uint32_t p6,
                                          uint32_t& top) {
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "madc.lo.cc.u32 %2, %10, %9, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %9, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %9, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %9, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %9, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(top)
        : "r"(m), "r"(p2), "r"(p4), "r"(p6));
}
// Y += m * (p1, p3, p5, p7) with p1 = 2^32 - 1: (Y1:Y0) += ((m - c) : t0)
__device__ __forceinline__ void redc_odd_p1allones(uint32_t* Y, uint32_t m, uint32_t t0, uint32_t p3, uint32_t p5,
                                                   uint32_t p7) {
    uint32_t c, hi;
    asm("min.u32 %0, %1, 1;" : "=r"(c) : "r"(t0));
    hi = m - c;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "madc.lo.cc.u32 %2, %11, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %8, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %8, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %8, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %8, %6;\n\t"
        "madc.hi.u32 %7, %13, %8, %7;"
        : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7])
        : "r"(m), "r"(t0), "r"(hi), "r"(p3), "r"(p5), "r"(p7));
}
// One operand-scanning row: (X | Y) += a * b ; then += m * p with m = -X[0] so that column 0 clears.
// X holds columns 0..7 ((0,1),(2,3),.. product pairs), Y holds columns 1..8.  After the row the
// roles swap (the caller alternates the arguments), which is the division by 2^32.
template <class F>
__device__ __forceinline__ void mont_row(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t b, bool first) {
    static_assert(F::p(0) == 1u, "the reduction rows assume p == 1 (mod 2^32)");
    if (first) {
        mul4(Y, a[1], a[3], a[5], a[7], b);
        mul4(X, a[0], a[2], a[4], a[6], b);
    } else {
        madc4_rshift(Y, X[0], a[1], a[3], a[5], a[7], b);
        cmad4(X, a[0], a[2], a[4], a[6], b, Y[7]);
    }
    // m = -X[0]: -p^-1 mod 2^32 == 0xffffffff for both moduli.  Kept opaque (asm volatile): if the
    // optimizer sees the negation it folds it into the multiply-adds below and ptxas then emits split
    // IMAD.X + IMAD.HI.U32.X pairs instead of one IMAD.WIDE.U32.X per product.
    const uint32_t t0 = X[0];
    uint32_t m;
    asm volatile("sub.u32 %0, 0, %1;" : "=r"(m) : "r"(t0));
    // total value stays < 2^288 (a < p), so the Y chain never carries out of column 8
    if (F::p(1) == 0xffffffffu)
        redc_odd_p1allones(Y, m, t0, F::p(3), F::p(5), F::p(7));
    else
        cmad4_nocarry(Y, F::p(1), F::p(3), F::p(5), F::p(7), m);
    redc_even(X, m, F::p(2), F::p(4), F::p(6), Y[7]);
}
}  // namespace detail

// Montgomery product a*b*R^-1 without the final conditional subtraction: a in [0,p), b in [0,2^256)
// -> result in [0,2p).
template <class F>
__device__ __forceinline__ Fe fe_mul_lazy(const Fe& a, const Fe& b) {
    uint32_t X[8], Y[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        detail::mont_row<F>(X, Y, a.v, b.v[i], i == 0);
        detail::mont_row<F>(Y, X, a.v, b.v[i + 1], false);
    }
    // eight rows done: X is back in the "columns 0.." role but its column 0 is the cleared one of
    // the last row's partner; the value is Y[1..7] (columns 1..7) + X[0..7] (columns 1..8), i.e. after
    // the last shift: r = X + (Y >> 32).
    Fe r;
    asm("add.cc.u32 %0,%8,%16;\n\taddc.cc.u32 %1,%9,%17;\n\taddc.cc.u32 %2,%10,%18;\n\taddc.cc.u32 %3,%11,%19;\n\t"
        "addc.cc.u32 %4,%12,%20;\n\taddc.cc.u32 %5,%13,%21;\n\taddc.cc.u32 %6,%14,%22;\n\taddc.u32 %7,%15,0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]), "r"(Y[1]), "r"(Y[2]),
          "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
    return r;
}

// Fully reduced Montgomery product (ark-ff `a * b`): a,b in [0,p) -> [0,p).
template <class F>
__device__ __forceinline__ Fe fe_mul(const Fe& a, const Fe& b) {
    return fe_reduce_once<F>(fe_mul_lazy<F>(a, b));
}


// ---- unreduced 512-bit product and deferred Montgomery reduction -------------------------------------
// out[0..15] = a * b as a plain integer (no reduction): the same even/odd carry-chain rows as fe_mul_lazy
// without the m*p half — 64 wide multiplies instead of 112.  Used for the LAST multiplication of each
// sumcheck term: sum_j (x_j * y_j) is accumulated unreduced (17 words) and reduced once per thread.
__device__ __forceinline__ void fe_mul_wide(uint32_t* out, const Fe& a, const Fe& b) {
    uint32_t X[8], Y[8];
    detail::mul4(Y, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
    detail::mul4(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
    out[0] = X[0];
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        // odd row: even-column role = Y, odd-column role = X
        detail::madc4_rshift(X, Y[0], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
        detail::cmad4(Y, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], X[7]);
        out[i] = Y[0];
        if (i + 1 < 8) {
            detail::madc4_rshift(Y, X[0], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
            detail::cmad4(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1], Y[7]);
            out[i + 1] = X[0];
        }
    }
    // after row 7 the even-column array is Y (Y[0] emitted, Y[1..7] = columns 1..7), X = columns 1..8
    asm("add.cc.u32 %0,%8,%16;\n\taddc.cc.u32 %1,%9,%17;\n\taddc.cc.u32 %2,%10,%18;\n\taddc.cc.u32 %3,%11,%19;\n\t"
        "addc.cc.u32 %4,%12,%20;\n\taddc.cc.u32 %5,%13,%21;\n\taddc.cc.u32 %6,%14,%22;\n\taddc.u32 %7,%15,0;"
        : "=r"(out[8]), "=r"(out[9]), "=r"(out[10]), "=r"(out[11]), "=r"(out[12]), "=r"(out[13]), "=r"(out[14]),
          "=r"(out[15])
        : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]), "r"(Y[1]), "r"(Y[2]),
          "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
}

// v[0..16] (a sum of < 2^22 such products, < 2^544) -> v * 2^-256 mod p, fully reduced.
// Nine reduction rows divide by 2^288 (the result is then < p + 2^256-ish/2^32 < 2p), one multiplication by
// 2^288 mod p restores the 2^-256 scaling of an ordinary Montgomery product.  Runs once per thread.
template <class F>
__device__ __noinline__ Fe fe_redc_wide(const uint32_t* vin) {
    uint32_t v[18];
#pragma unroll
    for (int i = 0; i < 17; i++) v[i] = vin[i];
    v[17] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint32_t m = 0u - v[i];
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)m * F::p(j) + v[i + j];
            v[i + j] = (uint32_t)c;
            c >>= 32;
        }
#pragma unroll
        for (int j = i + 8; j < 18; j++) {
            c += v[j];
            v[j] = (uint32_t)c;
            c >>= 32;
        }
    }
    Fe u, k;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u.v[i] = v[9 + i];
        k.v[i] = F::k288(i);
    }
    return fe_mul<F>(fe_reduce_once<F>(u), k);
}


// ---- multiplication by a launch-wide constant (the fold challenge r) ---------------------------------------
// Every fold of a round multiplies by the SAME r, so the host precomputes the eight multiples
//     tab.v[i] = r * 2^(32 i + 64) mod p          (plain integers, r canonical; host_field.hpp::fixed_mul_table)
// and x*r = sum_i x_i * tab[i] needs no interleaved reduction: eight ALIGNED rows (64 wide multiplies) give a
// sum < 2^290, two Montgomery rows divide the pre-scaled sum by 2^64 (result < 2p), one conditional subtract.
// 76 (Fr381) / 78 (Fr377) wide multiplies instead of 112 / 120.  For x = aR (Montgomery form) the result is the
// canonical residue (a r)R — bit-identical to fe_mul(x, rR).
namespace detail {
// One Montgomery row on U[0..8] (+U[9] if TOP): U += m*p with m = -U[0]; afterwards U[0] is (logically) zero.
template <class F, bool TOP>
__device__ __forceinline__ void redc_row_inplace(uint32_t* U) {
    const uint32_t t0 = U[0];
    uint32_t m;
    asm volatile("sub.u32 %0, 0, %1;" : "=r"(m) : "r"(t0));
    uint32_t top = TOP ? U[9] : 0u;
    // chain A: columns 1.. : the carry of column 0 (t0 + m*p0 = 2^32 * (t0 != 0)), m*p1 and the odd-column products
    if (F::p(1) == 0xffffffffu) {
        uint32_t b, s1, h2;
        asm("min.u32 %0, %1, 1;" : "=r"(b) : "r"(t0));
        const uint32_t mb = m - b;  // hi(m*p1); lo(m*p1) = t0
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(s1), "=r"(h2) : "r"(t0), "r"(b), "r"(mb));
        asm("add.cc.u32 %0, %0, %9;\n\t"
            "addc.cc.u32 %1, %1, %10;\n\t"
            "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
            "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
            "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
            "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
            "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
            "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(U[1]), "+r"(U[2]), "+r"(U[3]), "+r"(U[4]), "+r"(U[5]), "+r"(U[6]), "+r"(U[7]), "+r"(U[8]), "+r"(top)
            : "r"(s1), "r"(h2), "r"(F::p(3)), "r"(F::p(5)), "r"(F::p(7)), "r"(m));
    } else {
        uint32_t dead = t0;
        asm("add.cc.u32 %0, %0, %14;\n\t"
            "madc.lo.cc.u32 %1, %10, %14, %1;\n\t"
            "madc.hi.cc.u32 %2, %10, %14, %2;\n\t"
            "madc.lo.cc.u32 %3, %11, %14, %3;\n\t"
            "madc.hi.cc.u32 %4, %11, %14, %4;\n\t"
            "madc.lo.cc.u32 %5, %12, %14, %5;\n\t"
            "madc.hi.cc.u32 %6, %12, %14, %6;\n\t"
            "madc.lo.cc.u32 %7, %13, %14, %7;\n\t"
            "madc.hi.cc.u32 %8, %13, %14, %8;\n\t"
            "addc.u32 %9, %9, 0;"
            : "+r"(dead), "+r"(U[1]), "+r"(U[2]), "+r"(U[3]), "+r"(U[4]), "+r"(U[5]), "+r"(U[6]), "+r"(U[7]), "+r"(U[8]),
              "+r"(top)
            : "r"(F::p(1)), "r"(F::p(3)), "r"(F::p(5)), "r"(F::p(7)), "r"(m));
    }
    // chain B: the even-column products p2, p4, p6 at columns (2,3), (4,5), (6,7), then ripple
    asm("mad.lo.cc.u32 %0, %8, %11, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %11, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %11, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %11, %5;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(U[2]), "+r"(U[3]), "+r"(U[4]), "+r"(U[5]), "+r"(U[6]), "+r"(U[7]), "+r"(U[8]), "+r"(top)
        : "r"(F::p(2)), "r"(F::p(4)), "r"(F::p(6)), "r"(m));
    if (TOP) U[9] = top;
}
}  // namespace detail

template <class F>
__device__ __forceinline__ Fe fe_mul_fixed(const Fe& x, const FixedMul& tab) {
    uint32_t E[9], O[9];  // E[j] = column j, O[j] = column j+1; [8] collects the chain carries
    detail::mul4(E, tab.v[0][0], tab.v[0][2], tab.v[0][4], tab.v[0][6], x.v[0]);
    detail::mul4(O, tab.v[0][1], tab.v[0][3], tab.v[0][5], tab.v[0][7], x.v[0]);
    E[8] = 0;
    O[8] = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) {
        detail::cmad4(E, tab.v[i][0], tab.v[i][2], tab.v[i][4], tab.v[i][6], x.v[i], E[8]);
        detail::cmad4(O, tab.v[i][1], tab.v[i][3], tab.v[i][5], tab.v[i][7], x.v[i], O[8]);
    }
    uint32_t T[10];
    T[0] = E[0];
    asm("add.cc.u32 %0,%9,%17;\n\taddc.cc.u32 %1,%10,%18;\n\taddc.cc.u32 %2,%11,%19;\n\taddc.cc.u32 %3,%12,%20;\n\t"
        "addc.cc.u32 %4,%13,%21;\n\taddc.cc.u32 %5,%14,%22;\n\taddc.cc.u32 %6,%15,%23;\n\taddc.cc.u32 %7,%16,%24;\n\t"
        "addc.u32 %8,%25,0;"
        : "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(T[8]), "=r"(T[9])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(O[0]), "r"(O[1]),
          "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]));
    detail::redc_row_inplace<F, true>(T);       // columns 0..9 -> value in columns 1..9
    detail::redc_row_inplace<F, false>(T + 1);  // columns 1..9 -> value in columns 2..9, < 2p
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = T[2 + i];
    return fe_reduce_once<F>(r);
}

// fold with the precomputed multiples of the challenge: l - r*(l - h)
template <class F>
__device__ __forceinline__ Fe fe_fold_fixed(const Fe& l, const Fe& h, const FixedMul& tab) {
    return fe_sub<F>(l, fe_mul_fixed<F>(fe_sub<F>(l, h), tab));
}

// Montgomery form <-> canonical integer (ark-ff `into_bigint` / `from_bigint`)
template <class F>
__device__ __forceinline__ Fe fe_to_canonical(const Fe& a) {
    Fe one_raw = fe_zero<F>();
    one_raw.v[0] = 1;
    return fe_mul<F>(a, one_raw);
}
template <class F>
__device__ __forceinline__ Fe fe_from_canonical(const Fe& c) {  // c in [0,p)
    Fe r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.v[i] = F::r2(i);
    return fe_mul<F>(c, r2);
}

// a + t*(b - a) building block: fold(l, r, x) = l - x*(l - r)   (evaluation_form.rs:68)
template <class F>
__device__ __forceinline__ Fe fe_fold(const Fe& l, const Fe& r, const Fe& x) {
    return fe_sub<F>(l, fe_mul<F>(fe_sub<F>(l, r), x));
}

#endif  // __CUDACC__
}  // namespace zk
