// TEST INFRASTRUCTURE ONLY — host model of zk_b200/csrc/accw.cuh (+ fe_mul_wide / fe_redc_wide of field.cuh) for the
// thread-by-thread replays of sop_kernel.cuh.  Same layout and accessor arithmetic as the device code (so the shared
// memory carve-up is exercised), plain 64-bit carries instead of the PTX carry chains, and the Montgomery reduction
// of the 17-word sum done with host_field.hpp.  Include inside `namespace zk { namespace { ... } }` after kThreads,
// g_field, el() and fe() are defined.
inline void __syncthreads() {}

struct Accw {
    uint4* q;
    uint32_t* ov;
};
constexpr size_t accw_bytes(int np) { return (size_t)np * (4 * sizeof(uint4) + sizeof(uint32_t)) * kThreads; }
inline Accw accw_base(uint4* smem, int np) {
    return Accw{smem + threadIdx.x, reinterpret_cast<uint32_t*>(smem + (size_t)np * 4 * kThreads) + threadIdx.x};
}
inline Accw accw_at(const Accw& a, int t) { return Accw{a.q + t * 4 * kThreads, a.ov + t * kThreads}; }
inline void accw_zero(uint4* smem, int np) {
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = (int)threadIdx.x; i < (int)(accw_bytes(np) / 4); i += kThreads) w[i] = 0;
}
inline void accw_words(const Accw& a, uint32_t v[17], bool store) {
    for (int g = 0; g < 4; g++) {
        uint4& q = a.q[g * kThreads];
        uint32_t* p[4] = {&q.x, &q.y, &q.z, &q.w};
        for (int i = 0; i < 4; i++) {
            if (store) *p[i] = v[4 * g + i];
            else v[4 * g + i] = *p[i];
        }
    }
    if (store) a.ov[0] = v[16];
    else v[16] = a.ov[0];
}
inline void accw_add16(const Accw& a, const uint32_t* w) {
    uint32_t v[17];
    accw_words(a, v, false);
    uint64_t c = 0;
    for (int i = 0; i < 16; i++) { c += (uint64_t)v[i] + w[i]; v[i] = (uint32_t)c; c >>= 32; }
    v[16] += (uint32_t)c;
    accw_words(a, v, true);
}
inline void accw_add_hi(const Accw& a, const Fe& x) {
    uint32_t v[17];
    accw_words(a, v, false);
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)v[8 + i] + x.v[i]; v[8 + i] = (uint32_t)c; c >>= 32; }
    v[16] += (uint32_t)c;
    accw_words(a, v, true);
}
// out[0..15] = a * b as plain integers (the raw Montgomery residues)
inline void fe_mul_wide(uint32_t* out, const Fe& a, const Fe& b) {
    uint64_t t[17] = {0};
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.v[i] * b.v[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] += c;
    }
    for (int i = 0; i < 16; i++) out[i] = (uint32_t)t[i];
}
// T (17 words) -> T * 2^-256 mod p, fully reduced: T = sum_i w_i 2^(32 i) assembled in the Montgomery domain
// (S = T * R), then two Montgomery multiplications by the raw integer 1 (each multiplies by R^-1)
template <class F>
Fe accw_reduce(const Accw& a) {
    uint32_t v[17];
    accw_words(a, v, false);
    const host::Field& Fq = *g_field;
    const host::El two32 = Fq.from_u64((uint64_t)1 << 32);
    host::El s = Fq.zero(), pw = Fq.one();
    for (int i = 0; i < 17; i++) {
        s = Fq.add(s, Fq.mul(Fq.from_u64(v[i]), pw));
        pw = Fq.mul(pw, two32);
    }
    const host::El raw_one{{1, 0, 0, 0}};
    return fe(Fq.mul(Fq.mul(s, raw_one), raw_one));
}
