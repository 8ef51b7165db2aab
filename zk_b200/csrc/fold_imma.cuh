// fold_imma.cuh — the fold l + r (h - l) of a warp's 32 items by a launch-wide challenge r on the INT8 tensor path
// (mma.sync.m16n8k32.u8.u8.s32 -> IMMA.16832.U8.U8 on sm_100a).  Needs field.cuh.
//
// x = h - l (canonical, 32 bytes x_0 .. x_31) times r is sum_k x_k T_k with T_k = r * 2^(8 k + 32) mod p, a launch-wide
// table of 32 canonical integers.  Byte by byte that is a [32 items x 32] x [32 x 32] matrix product with 8-bit operands
// and 32-bit sums: column c of item i is S_c = sum_k x_k byte_c(T_k) <= 32 * 255^2 < 2^21, exact in the s32 accumulators,
// and x r 2^32 = sum_c S_c 2^(8 c) < 2^13 p.  Eight IMMA per warp and fold replace 128 DFMA + 64 uniform loads (or 76 wide
// multiplies) PER THREAD; what is left per thread is the re-assembly of the columns into limbs (ALU adds), ONE Montgomery
// row (6 wide multiplies) for the division by 2^32, a conditional subtract and the addition of l — the same field element
// as fe_fold_fixed / fe_fold_fixed_f64_x2 (both canonical), checked bit for bit by tools/imma_fold_probe.cu and by the parity suite.
//
// Operands change layout through a per-warp staging area in shared memory (kImmaStageBytes):
//   A: item-major rows of 32 B, the two 16-byte chunks swapped for rows 4..7 mod 8 (ldmatrix.x4 and the stores are then
//      conflict free).  Fragment (PTX ISA, m16n8k32 .u8): a0 = row g, bytes 4t..4t+3 = LIMB t of item g; a1 = row g + 8;
//      a2, a3 = the same rows, limb 4 + t   (g = lane >> 2, t = lane & 3)  == ldmatrix.x4 of the four 8 x 16-byte blocks.
//   B: b0 = bytes k = 4t..4t+3 of column n = g, b1 = k = 16 + 4t..: eight registers per lane hold the whole table
//      (host_fixed_mul_table_i8 stores it in fragment order).
//   C: c0, c1 = row g, columns 2t, 2t+1; c2, c3 = row g + 8: the pair c0 + (c1 << 8) < 2^30 is the 16-bit-aligned word
//      4 nt + t of the item's sum; rows of 16 such words (the 16-byte chunks XOR-swizzled by (row >> 1) & 3: conflict free
//      both ways) are read back by the item's own lane.
#pragma once
#include <cstdint>

#include "host_field.hpp"

namespace zk {

constexpr int kImmaStageBytes = 32 * 64;
struct FixedMulI8 {
    uint32_t frag[32][8];  // [lane][2 * nt + half]
};
struct ImmaTab {
    uint32_t b[8];
};

#ifdef __CUDACC__
__device__ __forceinline__ ImmaTab imma_tab_load(const FixedMulI8& tab, int lane) {
    ImmaTab t;
#pragma unroll
    for (int k = 0; k < 8; k++) t.b[k] = tab.frag[lane][k];
    return t;
}

// NF folds at once: out[f] = l[f] + r (h[f] - l[f]).  ALL 32 lanes of the warp must call (mma.sync, __syncwarp); `stage` is
// the warp's own NF * kImmaStageBytes, 16-byte aligned.  NF = 2 (the two folds of a fused-round item) halves the number of
// warp synchronisations and gives the scheduler sixteen independent IMMA.
template <class F, int NF>
__device__ __forceinline__ void fe_fold_imma_n(Fe* out, const Fe* l, const Fe* h, const ImmaTab& tab, unsigned char* stage, int lane) {
#pragma unroll
    for (int f = 0; f < NF; f++) {
        const Fe x = fe_sub<F>(h[f], l[f]);
        uint4* row = reinterpret_cast<uint4*>(stage + f * kImmaStageBytes + lane * 32);
        const int sw = (lane >> 2) & 1;
        row[sw] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
        row[sw ^ 1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    }
    __syncwarp();
    uint32_t a[NF][2][4];
    {
        const int mi = lane >> 3;  // which 8 x 16-byte block this lane addresses
        const int item = (mi & 1) * 8 + (lane & 7), chunk = (mi >> 1) ^ ((item >> 2) & 1);
#pragma unroll
        for (int f = 0; f < NF; f++)
#pragma unroll
            for (int T = 0; T < 2; T++) {
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(stage + f * kImmaStageBytes + (16 * T + item) * 32 + chunk * 16);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a[f][T][0]), "=r"(a[f][T][1]), "=r"(a[f][T][2]), "=r"(a[f][T][3])
                             : "r"(addr));
            }
    }
    __syncwarp();  // every fragment load is done: the staging area is reused for the sums
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int f = 0; f < NF; f++) {
        uint32_t* const words = reinterpret_cast<uint32_t*>(stage + f * kImmaStageBytes);
#pragma unroll
        for (int T = 0; T < 2; T++) {
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3)
                             : "r"(a[f][T][0]), "r"(a[f][T][1]), "r"(a[f][T][2]), "r"(a[f][T][3]), "r"(tab.b[2 * nt]), "r"(tab.b[2 * nt + 1]));
                // rows 16 T + g and 16 T + g + 8 share the swizzle ((row >> 1) & 3 == (g >> 1) & 3)
                const int pc = (nt ^ ((g >> 1) & 3)) * 4 + t;
                words[(16 * T + g) * 16 + pc] = (uint32_t)c0 + ((uint32_t)c1 << 8);
                words[(16 * T + g + 8) * 16 + pc] = (uint32_t)c2 + ((uint32_t)c3 << 8);
            }
        }
    }
    __syncwarp();
    uint32_t w[NF][16];
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 v = *reinterpret_cast<const uint4*>(stage + f * kImmaStageBytes + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
            w[f][4 * q] = v.x;
            w[f][4 * q + 1] = v.y;
            w[f][4 * q + 2] = v.z;
            w[f][4 * q + 3] = v.w;
        }
    __syncwarp();  // the next call's staging stores must not overtake these loads
#pragma unroll
    for (int f = 0; f < NF; f++) {
        // sum_k w_k 2^(16 k) as nine 32-bit limbs: the even words are limbs as they stand, the odd words are limbs of a
        // number shifted up by 16 bits (funnel shifts and one carry chain on the ALU pipe; written as 64-bit arithmetic
        // ptxas turns it into wide multiplies by 65536)
        uint32_t V[9], s[9];
        s[0] = w[f][1] << 16;
#pragma unroll
        for (int j = 1; j < 8; j++) s[j] = __funnelshift_l(w[f][2 * j - 1], w[f][2 * j + 1], 16);
        s[8] = w[f][15] >> 16;
        asm("add.cc.u32 %0,%9,%18;\n\taddc.cc.u32 %1,%10,%19;\n\taddc.cc.u32 %2,%11,%20;\n\taddc.cc.u32 %3,%12,%21;\n\t"
            "addc.cc.u32 %4,%13,%22;\n\taddc.cc.u32 %5,%14,%23;\n\taddc.cc.u32 %6,%15,%24;\n\taddc.cc.u32 %7,%16,%25;\n\t"
            "addc.u32 %8,%17,0;"
            : "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), "=r"(V[8])
            : "r"(w[f][0]), "r"(w[f][2]), "r"(w[f][4]), "r"(w[f][6]), "r"(w[f][8]), "r"(w[f][10]), "r"(w[f][12]), "r"(w[f][14]), "r"(s[8]),
              "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]));
        detail::redc_row_inplace<F, false>(V);  // columns 0..8 -> value in columns 1..8, < 2p
        Fe rx;
#pragma unroll
        for (int i = 0; i < 8; i++) rx.v[i] = V[1 + i];
        out[f] = fe_add<F>(l[f], fe_reduce_once<F>(rx));
    }
}
template <class F>
__device__ __forceinline__ Fe fe_fold_imma(const Fe& l, const Fe& h, const ImmaTab& tab, unsigned char* stage, int lane) {
    Fe out;
    fe_fold_imma_n<F, 1>(&out, &l, &h, tab, stage, lane);
    return out;
}
#endif  // __CUDACC__

// Host side: T_k = r * 2^(8 k + 32) mod p, canonical, in B-fragment order.  `r_mont` is r in
// Montgomery form (as for fixed_mul_table).
inline void fixed_mul_table_i8(const host::Field& F, const host::El& r_mont, FixedMulI8* out) {
    unsigned char T[32][32];
    const host::El two8 = F.from_u64(256);
    host::El cur = F.mul(r_mont, F.from_u64((uint64_t)1 << 32));
    for (int k = 0; k < 32; k++) {
        uint64_t c[4];
        F.to_canonical(cur, c);
        for (int b = 0; b < 32; b++) T[k][b] = (unsigned char)(c[b >> 3] >> (8 * (b & 7)));
        cur = F.mul(cur, two8);
    }
    for (int lane = 0; lane < 32; lane++) {
        const int g = lane >> 2, t = lane & 3;
        for (int nt = 0; nt < 4; nt++)
            for (int half = 0; half < 2; half++) {
                uint32_t v = 0;
                for (int b = 0; b < 4; b++) v |= (uint32_t)T[16 * half + 4 * t + b][8 * nt + g] << (8 * b);
                out->frag[lane][2 * nt + half] = v;
            }
    }
}

}  // namespace zk
