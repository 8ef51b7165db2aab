// field29.cuh — carry-free Montgomery multiplication in a reduced radix (9 limbs x 29 bits, R' = 2^261).
//
// NEGATIVE RESULT, kept for the record (used only by tools/ubench.cu).  Hypothesis: IMAD.WIDE.U32 with carry
// predicates is slower than a plain IMAD.WIDE.U32, so a radix-2^29 representation whose columns accumulate in one
// 64-bit register pair without any carry should win.  Measured on B200: carries are free (29.7 vs 30.5 wide
// multiplies per clock per SM) and the wide multiply itself is the half-rate, binding instruction; this form
// needs 153-162 of them (9x9 + 9x8 + 9) against 112 for the radix-2^32 multiplier in field.cuh, and ptxas
// additionally rewrites the accumulation into IMAD.WIDE(..,RZ) + 3-input IADD3 trees.  0.165 vs 0.25 mul/clk/SM.
// It is bit-exact (mul29(a, 32*b) == fe_mul(a, b) on 2^20 random pairs, both fields).
//     column k :  C += sum_{i+j=k} a_i b_j  +  sum_{i+j=k, j>0} m_i p_j ;   m_k = -C mod 2^29 ;  C = (C + m_k) >> 29
// Scaling: mul29(x, y) = x*y*2^-261 mod p.
#pragma once
#include "../zk_b200/csrc/field.cuh"

namespace zk {

struct Fe29 {
    uint32_t l[9];
};

constexpr uint32_t kMask29 = (1u << 29) - 1;

template <class F>
struct P29;
template <>
struct P29<Fr381> {
    __host__ __device__ static constexpr uint32_t p(int k) {
        constexpr uint32_t t[9] = {0x00000001u, 0x1ffffff8u, 0x1f96ffbfu, 0x1b4805ffu, 0x1d80553bu,
                                   0x0c0404d0u, 0x1520cce7u, 0x0a6533afu, 0x0073eda7u};
        return t[k];
    }
};
template <>
struct P29<Fr377> {
    __host__ __device__ static constexpr uint32_t p(int k) {
        constexpr uint32_t t[9] = {0x00000001u, 0x108c0000u, 0x00000042u, 0x14edfda0u, 0x1b00159au,
                                   0x068f2e1bu, 0x155982d1u, 0x0bd34594u, 0x0012ab65u};
        return t[k];
    }
};

#ifdef __CUDACC__

// 8 x 32-bit words -> 9 x 29-bit limbs (pure bit regrouping; the value is unchanged)
__device__ __forceinline__ Fe29 unpack29(const Fe& w) {
    Fe29 r;
    r.l[0] = w.v[0] & kMask29;
    r.l[1] = __funnelshift_r(w.v[0], w.v[1], 29) & kMask29;
    r.l[2] = __funnelshift_r(w.v[1], w.v[2], 26) & kMask29;
    r.l[3] = __funnelshift_r(w.v[2], w.v[3], 23) & kMask29;
    r.l[4] = __funnelshift_r(w.v[3], w.v[4], 20) & kMask29;
    r.l[5] = __funnelshift_r(w.v[4], w.v[5], 17) & kMask29;
    r.l[6] = __funnelshift_r(w.v[5], w.v[6], 14) & kMask29;
    r.l[7] = __funnelshift_r(w.v[6], w.v[7], 11) & kMask29;
    r.l[8] = w.v[7] >> 8;
    return r;
}
// 9 normalised limbs (each < 2^29, value < 2^256) -> 8 words
__device__ __forceinline__ Fe pack29(const Fe29& a) {
    Fe r;
    r.v[0] = a.l[0] | (a.l[1] << 29);
    r.v[1] = __funnelshift_r(a.l[1] << 3, a.l[2], 6);
    r.v[2] = __funnelshift_r(a.l[2] << 3, a.l[3], 9);
    r.v[3] = __funnelshift_r(a.l[3] << 3, a.l[4], 12);
    r.v[4] = __funnelshift_r(a.l[4] << 3, a.l[5], 15);
    r.v[5] = __funnelshift_r(a.l[5] << 3, a.l[6], 18);
    r.v[6] = __funnelshift_r(a.l[6] << 3, a.l[7], 21);
    r.v[7] = __funnelshift_r(a.l[7] << 3, a.l[8], 24);
    return r;
}

// Montgomery product x*y*2^-261 mod p, carry free.
//   a : limbs < 2^31 (may be an un-normalised sum/difference), b : limbs < 2^29.
//   result: limbs < 2^29 (top limb < 2^27), value < a*b/2^261 + p.
template <class F>
__device__ __forceinline__ Fe29 mul29(const Fe29& a, const Fe29& b) {
    uint64_t C = 0;
    uint32_t m[9];
    Fe29 r;
#pragma unroll
    for (int k = 0; k < 9; k++) {
#pragma unroll
        for (int i = 0; i <= k; i++) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(C) : "r"(a.l[i]), "r"(b.l[k - i]));
#pragma unroll
        for (int i = 0; i < k; i++) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(C) : "r"(m[i]), "r"(P29<F>::p(k - i)));
        m[k] = (0u - (uint32_t)C) & kMask29;  // -p^-1 == -1 (mod 2^29), p_0 == 1
        asm("mad.wide.u32 %0, %1, 1, %0;" : "+l"(C) : "r"(m[k]));
        C >>= 29;
    }
#pragma unroll
    for (int k = 9; k < 17; k++) {
#pragma unroll
        for (int i = k - 8; i <= 8; i++) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(C) : "r"(a.l[i]), "r"(b.l[k - i]));
#pragma unroll
        for (int i = k - 8; i <= 8; i++) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(C) : "r"(m[i]), "r"(P29<F>::p(k - i)));
        r.l[k - 9] = (uint32_t)C & kMask29;
        C >>= 29;
    }
    r.l[8] = (uint32_t)C;
    return r;
}

// limb-wise lazy add: limbs grow by at most one bit
__device__ __forceinline__ Fe29 add29(const Fe29& a, const Fe29& b) {
    Fe29 r;
#pragma unroll
    for (int k = 0; k < 9; k++) r.l[k] = a.l[k] + b.l[k];
    return r;
}
// carry propagation: limbs < 2^32 -> limbs < 2^29 (top limb takes the rest)
__device__ __forceinline__ Fe29 norm29(const Fe29& a) {
    Fe29 r;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t t = a.l[k] + c;
        r.l[k] = t & kMask29;
        c = t >> 29;
    }
    r.l[8] = a.l[8] + c;
    return r;
}

#endif  // __CUDACC__
}  // namespace zk
