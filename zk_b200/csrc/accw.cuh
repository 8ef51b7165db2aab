// accw.cuh — wide (17-word) per-thread accumulators in shared memory: the deferred Montgomery reduction of the round
// kernels (kernels_sumcheck.cu, the WIDE variant of sop_kernel.cuh).  The LAST multiplication of a term is a plain
// 256x256 -> 512-bit product (fe_mul_wide) added here; one reduction per thread and evaluation point at the end
// (sum of products then one REDC == sum of REDCs, exactly, mod p).  Device only (inline PTX carry chains); needs
// kThreads (reduce.cuh) and field.cuh.
#pragma once

namespace zk {
namespace {

// ---- wide (17-word) per-thread accumulators in shared memory -------------------------------------------
// Layout per evaluation point t: uint4 q[4][kThreads] (words 0..15 = the unreduced sum) and uint32_t ov[kThreads]
// (word 16 = overflow count).  An Accw points at this thread's column of point 0; point t is `at(t, D+1)`.
struct Accw {
    uint4* q;       // + g * kThreads, g = 0..3
    uint32_t* ov;
};
__host__ __device__ constexpr size_t accw_bytes(int np) { return (size_t)np * (4 * sizeof(uint4) + sizeof(uint32_t)) * kThreads; }
__device__ __forceinline__ Accw accw_base(uint4* smem, int np) {
    return Accw{smem + threadIdx.x, reinterpret_cast<uint32_t*>(smem + (size_t)np * 4 * kThreads) + threadIdx.x};
}
__device__ __forceinline__ Accw accw_at(const Accw& a, int t) { return Accw{a.q + t * 4 * kThreads, a.ov + t * kThreads}; }
__device__ __forceinline__ void accw_zero(uint4* smem, int np) {
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < (int)(accw_bytes(np) / 4); i += kThreads) w[i] = 0;
}
// acc += w[0..15]  (one 17-word carry chain)
__device__ __forceinline__ void accw_add16(const Accw& a, const uint32_t* w) {
    uint4 q0 = a.q[0 * kThreads], q1 = a.q[1 * kThreads], q2 = a.q[2 * kThreads], q3 = a.q[3 * kThreads];
    uint32_t ov = a.ov[0];
    asm("add.cc.u32 %0,%0,%17;\n\taddc.cc.u32 %1,%1,%18;\n\taddc.cc.u32 %2,%2,%19;\n\taddc.cc.u32 %3,%3,%20;\n\t"
        "addc.cc.u32 %4,%4,%21;\n\taddc.cc.u32 %5,%5,%22;\n\taddc.cc.u32 %6,%6,%23;\n\taddc.cc.u32 %7,%7,%24;\n\t"
        "addc.cc.u32 %8,%8,%25;\n\taddc.cc.u32 %9,%9,%26;\n\taddc.cc.u32 %10,%10,%27;\n\taddc.cc.u32 %11,%11,%28;\n\t"
        "addc.cc.u32 %12,%12,%29;\n\taddc.cc.u32 %13,%13,%30;\n\taddc.cc.u32 %14,%14,%31;\n\taddc.cc.u32 %15,%15,%32;\n\t"
        "addc.u32 %16,%16,0;"
        : "+r"(q0.x), "+r"(q0.y), "+r"(q0.z), "+r"(q0.w), "+r"(q1.x), "+r"(q1.y), "+r"(q1.z), "+r"(q1.w), "+r"(q2.x),
          "+r"(q2.y), "+r"(q2.z), "+r"(q2.w), "+r"(q3.x), "+r"(q3.y), "+r"(q3.z), "+r"(q3.w), "+r"(ov)
        : "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
          "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]));
    a.q[0 * kThreads] = q0; a.q[1 * kThreads] = q1; a.q[2 * kThreads] = q2; a.q[3 * kThreads] = q3;
    a.ov[0] = ov;
}
// acc += x * 2^256 (x a reduced element): what an ordinary product x*R contributes before the deferred reduction
__device__ __forceinline__ void accw_add_hi(const Accw& a, const Fe& x) {
    uint4 q2 = a.q[2 * kThreads], q3 = a.q[3 * kThreads];
    uint32_t ov = a.ov[0];
    asm("add.cc.u32 %0,%0,%9;\n\taddc.cc.u32 %1,%1,%10;\n\taddc.cc.u32 %2,%2,%11;\n\taddc.cc.u32 %3,%3,%12;\n\t"
        "addc.cc.u32 %4,%4,%13;\n\taddc.cc.u32 %5,%5,%14;\n\taddc.cc.u32 %6,%6,%15;\n\taddc.cc.u32 %7,%7,%16;\n\t"
        "addc.u32 %8,%8,0;"
        : "+r"(q2.x), "+r"(q2.y), "+r"(q2.z), "+r"(q2.w), "+r"(q3.x), "+r"(q3.y), "+r"(q3.z), "+r"(q3.w), "+r"(ov)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]));
    a.q[2 * kThreads] = q2; a.q[3 * kThreads] = q3;
    a.ov[0] = ov;
}
template <class F>
__device__ __forceinline__ Fe accw_reduce(const Accw& a) {
    uint32_t v[17];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint4 q = a.q[i * kThreads];
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
    v[16] = a.ov[0];
    return fe_redc_wide<F>(v);
}

}  // namespace
}  // namespace zk
