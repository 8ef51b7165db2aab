// keccak_avx512.cpp — Keccak-f[1600] absorb loop with AVX-512 (host; compiled by g++ with -mavx512f -mavx512vl).
//
// `SumcheckProver::prove` and `SumcheckVerifier::verify` absorb `poly.to_bytes()` — every table entry, 32 bytes each —
// into the transcript before the first round (sumcheck/src/prover.rs:16-17, verifier.rs:22-23).  The sponge is
// sequential by construction, so at 2^20+ entries that single host thread is the whole cost of the non-partial API
// (1 GiB at BASELINE config 2).  The portable permutation in keccak.hpp runs at about 8 cycles/byte; this one keeps the
// state in five zmm registers across blocks and needs 21 shuffle-port operations per round (about 4 cycles/byte).
//
// Layout L: register P[y], lane x (0..4) = lane A[x][y] of the state (lanes 5..7 unused).  One round:
//   theta  C = P0^..^P4 (two ternary-logic ops), D = perm(C, x-1) ^ rol(perm(C, x+1), 1)
//   rho    E[y] = rolv(P[y] ^ D, r[.][y])                       (per-lane variable rotate)
//   pi     B[y'][x'] = E[y][x] with (x', y') = (y, 2x+3y): F[y] = perm(E[y]) puts the element bound for plane y' in
//          lane y', so that F[p] lane l = B[l][p] — the state transposed
//   chi    G[p] = F[p] ^ (~F[p+1] & F[p+2]) is then a purely vertical ternary-logic op; iota on G[0] lane 0
//   back   P[y] lane x = G[x] lane y: a 5x5 transpose (4 unpacks + 10 two-source permutes)
// Dispatch: keccak.hpp calls zk_keccak256_absorb_avx512 when zk_keccak_avx512_available() (CPUID) says so; the
// CPU suite checks it against the portable permutation on random inputs of every length class.
#include <immintrin.h>

#include <cstddef>
#include <cstdint>
#include <cstdlib>

namespace {

alignas(64) const uint64_t kRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

// rho offsets r[x][y], one vector per plane y (lane x)
alignas(64) const uint64_t kRho[5][8] = {{0, 1, 62, 28, 27, 0, 0, 0},
                                         {36, 44, 6, 55, 20, 0, 0, 0},
                                         {3, 10, 43, 25, 39, 0, 0, 0},
                                         {41, 45, 15, 21, 8, 0, 0, 0},
                                         {18, 2, 61, 56, 14, 0, 0, 0}};
// pi: F[y] lane y' = E[y] lane (3 y' + y) mod 5
alignas(64) const uint64_t kPi[5][8] = {{0, 3, 1, 4, 2, 5, 6, 7},
                                        {1, 4, 2, 0, 3, 5, 6, 7},
                                        {2, 0, 3, 1, 4, 5, 6, 7},
                                        {3, 1, 4, 2, 0, 5, 6, 7},
                                        {4, 2, 0, 3, 1, 5, 6, 7}};
alignas(64) const uint64_t kRotL[8] = {4, 0, 1, 2, 3, 5, 6, 7};  // lane x <- lane x-1
alignas(64) const uint64_t kRotR[8] = {1, 2, 3, 4, 0, 5, 6, 7};  // lane x <- lane x+1
// transpose: with L01 = unpacklo(G0,G1) = (G0[0],G1[0],G0[2],G1[2],G0[4],G1[4],..), H01 = unpackhi = (G0[1],G1[1],G0[3],G1[3],..)
// and the same for (G2,G3):  Q[y] = (pair_y of *01, pair_y of *23) via one two-source permute, then lane 4 <- G4[y].
alignas(64) const uint64_t kPair[3][8] = {{0, 1, 8, 9, 4, 5, 6, 7},     // pair in lanes 0,1 (y = 0 from L, y = 1 from H)
                                          {2, 3, 10, 11, 4, 5, 6, 7},   // pair in lanes 2,3 (y = 2 from L, y = 3 from H)
                                          {4, 5, 12, 13, 4, 5, 6, 7}};  // pair in lanes 4,5 (y = 4 from L)
alignas(64) const uint64_t kLast[5][8] = {{0, 1, 2, 3, 8, 5, 6, 7},
                                          {0, 1, 2, 3, 9, 5, 6, 7},
                                          {0, 1, 2, 3, 10, 5, 6, 7},
                                          {0, 1, 2, 3, 11, 5, 6, 7},
                                          {0, 1, 2, 3, 12, 5, 6, 7}};

#define LD(p) _mm512_load_si512((const void*)(p))
#define XOR3(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0x96)
#define CHI(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0xD2)  // a ^ (~b & c)

struct State {
    __m512i p0, p1, p2, p3, p4;
};

__attribute__((always_inline)) inline void permute(State& s) {
    const __m512i rotl = LD(kRotL), rotr = LD(kRotR);
    const __m512i rho0 = LD(kRho[0]), rho1 = LD(kRho[1]), rho2 = LD(kRho[2]), rho3 = LD(kRho[3]), rho4 = LD(kRho[4]);
    const __m512i pi0 = LD(kPi[0]), pi1 = LD(kPi[1]), pi2 = LD(kPi[2]), pi3 = LD(kPi[3]), pi4 = LD(kPi[4]);
    const __m512i pr0 = LD(kPair[0]), pr1 = LD(kPair[1]), pr2 = LD(kPair[2]);
    const __m512i l0 = LD(kLast[0]), l1 = LD(kLast[1]), l2 = LD(kLast[2]), l3 = LD(kLast[3]), l4 = LD(kLast[4]);
    __m512i p0 = s.p0, p1 = s.p1, p2 = s.p2, p3 = s.p3, p4 = s.p4;
    for (int r = 0; r < 24; r++) {
        // theta
        const __m512i c = XOR3(XOR3(p0, p1, p2), p3, p4);
        const __m512i d = _mm512_xor_si512(_mm512_permutexvar_epi64(rotl, c), _mm512_rol_epi64(_mm512_permutexvar_epi64(rotr, c), 1));
        // rho, then the lane half of pi
        const __m512i f0 = _mm512_permutexvar_epi64(pi0, _mm512_rolv_epi64(_mm512_xor_si512(p0, d), rho0));
        const __m512i f1 = _mm512_permutexvar_epi64(pi1, _mm512_rolv_epi64(_mm512_xor_si512(p1, d), rho1));
        const __m512i f2 = _mm512_permutexvar_epi64(pi2, _mm512_rolv_epi64(_mm512_xor_si512(p2, d), rho2));
        const __m512i f3 = _mm512_permutexvar_epi64(pi3, _mm512_rolv_epi64(_mm512_xor_si512(p3, d), rho3));
        const __m512i f4 = _mm512_permutexvar_epi64(pi4, _mm512_rolv_epi64(_mm512_xor_si512(p4, d), rho4));
        // chi (vertical in the transposed layout) + iota
        __m512i g0 = CHI(f0, f1, f2);
        const __m512i g1 = CHI(f1, f2, f3), g2 = CHI(f2, f3, f4), g3 = CHI(f3, f4, f0), g4 = CHI(f4, f0, f1);
        g0 = _mm512_xor_si512(g0, _mm512_maskz_set1_epi64(0x01, (long long)kRC[r]));
        // transpose back: P[y] lane x = G[x] lane y
        const __m512i L01 = _mm512_unpacklo_epi64(g0, g1), H01 = _mm512_unpackhi_epi64(g0, g1);
        const __m512i L23 = _mm512_unpacklo_epi64(g2, g3), H23 = _mm512_unpackhi_epi64(g2, g3);
        p0 = _mm512_permutex2var_epi64(_mm512_permutex2var_epi64(L01, pr0, L23), l0, g4);
        p1 = _mm512_permutex2var_epi64(_mm512_permutex2var_epi64(H01, pr0, H23), l1, g4);
        p2 = _mm512_permutex2var_epi64(_mm512_permutex2var_epi64(L01, pr1, L23), l2, g4);
        p3 = _mm512_permutex2var_epi64(_mm512_permutex2var_epi64(H01, pr1, H23), l3, g4);
        p4 = _mm512_permutex2var_epi64(_mm512_permutex2var_epi64(L01, pr2, L23), l4, g4);
    }
    s.p0 = p0; s.p1 = p1; s.p2 = p2; s.p3 = p3; s.p4 = p4;
}

}  // namespace

extern "C" {

// 1 when the CPU executes AVX-512F/VL (the OS-enabled state is part of the builtin's check)
__attribute__((visibility("hidden"))) int zk_keccak_avx512_available() {
    static const int ok = [] {
        const char* e = std::getenv("ZK_B200_KECCAK");  // "portable" pins the scalar permutation (A/B measurements)
        if (e && e[0] == 'p') return 0;
        return (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl")) ? 1 : 0;
    }();
    return ok;
}

// state[25] (lane x + 5y, the layout of keccak.hpp) absorbs `nblocks` rate-136 blocks: xor 17 words, permute.
__attribute__((visibility("hidden"))) void zk_keccak256_absorb_avx512(uint64_t* state, const uint8_t* data, size_t nblocks) {
    State s;
    s.p0 = _mm512_maskz_loadu_epi64(0x1F, state + 0);
    s.p1 = _mm512_maskz_loadu_epi64(0x1F, state + 5);
    s.p2 = _mm512_maskz_loadu_epi64(0x1F, state + 10);
    s.p3 = _mm512_maskz_loadu_epi64(0x1F, state + 15);
    s.p4 = _mm512_maskz_loadu_epi64(0x1F, state + 20);
    for (size_t b = 0; b < nblocks; b++, data += 136) {
        s.p0 = _mm512_xor_si512(s.p0, _mm512_maskz_loadu_epi64(0x1F, data + 0));
        s.p1 = _mm512_xor_si512(s.p1, _mm512_maskz_loadu_epi64(0x1F, data + 40));
        s.p2 = _mm512_xor_si512(s.p2, _mm512_maskz_loadu_epi64(0x1F, data + 80));
        s.p3 = _mm512_xor_si512(s.p3, _mm512_maskz_loadu_epi64(0x03, data + 120));
        permute(s);
    }
    _mm512_mask_storeu_epi64(state + 0, 0x1F, s.p0);
    _mm512_mask_storeu_epi64(state + 5, 0x1F, s.p1);
    _mm512_mask_storeu_epi64(state + 10, 0x1F, s.p2);
    _mm512_mask_storeu_epi64(state + 15, 0x1F, s.p3);
    _mm512_mask_storeu_epi64(state + 20, 0x1F, s.p4);
}

}  // extern "C"
