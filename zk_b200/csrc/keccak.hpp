// keccak.hpp — host Keccak-256 and the Fiat-Shamir transcript (transcript/src/lib.rs:5-35).
// The reference uses sha3 0.10.8 `Keccak256` (original Keccak padding 0x01, rate 136), through
// `update` and `finalize_reset`.  Hashing stays on the host by design (BASELINE.json north_star).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "host_field.hpp"

// keccak_avx512.cpp: the bulk absorb loop with the state held in five zmm registers (about twice the portable speed)
extern "C" int zk_keccak_avx512_available();
extern "C" void zk_keccak256_absorb_avx512(uint64_t* state, const uint8_t* data, size_t nblocks);

namespace zk {
namespace host {

class Keccak256 {
   public:
    // allow_simd = false pins the portable permutation (the differential test of the AVX-512 path)
    explicit Keccak256(bool allow_simd = true) : simd_(allow_simd && zk_keccak_avx512_available() != 0) { reset(); }
    void reset() { std::memset(s_, 0, sizeof s_); fill_ = 0; }
    void update(const uint8_t* d, size_t n) {
        if (fill_) {
            size_t take = kRate - fill_;
            if (take > n) take = n;
            std::memcpy(buf_ + fill_, d, take);
            fill_ += take; d += take; n -= take;
            if (fill_ == kRate) { absorb(buf_); fill_ = 0; }
        }
        if (simd_ && n >= kRate) {  // whole blocks: the table absorb of prove()/verify() lives here
            const size_t nb = n / kRate;
            zk_keccak256_absorb_avx512(s_, d, nb);
            d += nb * kRate; n -= nb * kRate;
        }
        while (n >= kRate) { absorb(d); d += kRate; n -= kRate; }
        if (n) { std::memcpy(buf_, d, n); fill_ = n; }
    }
    void finalize_reset(uint8_t out[32]) {
        std::memset(buf_ + fill_, 0, kRate - fill_);
        buf_[fill_] ^= 0x01;
        buf_[kRate - 1] ^= 0x80;
        absorb(buf_);
        std::memcpy(out, s_, 32);
        reset();
    }

   private:
    static constexpr size_t kRate = 136;
    static inline uint64_t rol(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
    void absorb(const uint8_t* b) {
        for (int i = 0; i < 17; i++) { uint64_t l; std::memcpy(&l, b + 8 * i, 8); s_[i] ^= l; }
        permute();
    }
    // Keccak-f[1600], fully unrolled per round with the state in locals (the absorb of prove()/verify()
    // hashes the whole table sequentially on one host thread: this loop is that path's floor).
    void permute() {
        static const uint64_t RC[24] = {
            0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
            0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
            0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
            0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
            0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
            0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
        uint64_t a00 = s_[0], a01 = s_[1], a02 = s_[2], a03 = s_[3], a04 = s_[4], a05 = s_[5], a06 = s_[6], a07 = s_[7],
                 a08 = s_[8], a09 = s_[9], a10 = s_[10], a11 = s_[11], a12 = s_[12], a13 = s_[13], a14 = s_[14],
                 a15 = s_[15], a16 = s_[16], a17 = s_[17], a18 = s_[18], a19 = s_[19], a20 = s_[20], a21 = s_[21],
                 a22 = s_[22], a23 = s_[23], a24 = s_[24];
        for (int r = 0; r < 24; r++) {
            // theta
            const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21,
                           c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22, c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23,
                           c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
            const uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1),
                           d4 = c3 ^ rol(c0, 1);
            // rho + pi: B[y][2x+3y] = rol(A[x][y] ^ D[x], r[x][y]), lanes indexed x + 5y
            const uint64_t b00 = a00 ^ d0, b10 = rol(a01 ^ d1, 1), b20 = rol(a02 ^ d2, 62), b05 = rol(a03 ^ d3, 28),
                           b15 = rol(a04 ^ d4, 27), b16 = rol(a05 ^ d0, 36), b01 = rol(a06 ^ d1, 44),
                           b11 = rol(a07 ^ d2, 6), b21 = rol(a08 ^ d3, 55), b06 = rol(a09 ^ d4, 20),
                           b07 = rol(a10 ^ d0, 3), b17 = rol(a11 ^ d1, 10), b02 = rol(a12 ^ d2, 43),
                           b12 = rol(a13 ^ d3, 25), b22 = rol(a14 ^ d4, 39), b23 = rol(a15 ^ d0, 41),
                           b08 = rol(a16 ^ d1, 45), b18 = rol(a17 ^ d2, 15), b03 = rol(a18 ^ d3, 21),
                           b13 = rol(a19 ^ d4, 8), b14 = rol(a20 ^ d0, 18), b24 = rol(a21 ^ d1, 2),
                           b09 = rol(a22 ^ d2, 61), b19 = rol(a23 ^ d3, 56), b04 = rol(a24 ^ d4, 14);
            // chi (+ iota on lane 0)
            a00 = b00 ^ (~b01 & b02) ^ RC[r]; a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04);
            a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
            a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09);
            a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
            a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14);
            a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
            a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19);
            a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
            a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24);
            a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
        }
        s_[0] = a00; s_[1] = a01; s_[2] = a02; s_[3] = a03; s_[4] = a04; s_[5] = a05; s_[6] = a06; s_[7] = a07;
        s_[8] = a08; s_[9] = a09; s_[10] = a10; s_[11] = a11; s_[12] = a12; s_[13] = a13; s_[14] = a14; s_[15] = a15;
        s_[16] = a16; s_[17] = a17; s_[18] = a18; s_[19] = a19; s_[20] = a20; s_[21] = a21; s_[22] = a22; s_[23] = a23;
        s_[24] = a24;
    }
    uint64_t s_[25];
    uint8_t buf_[kRate];
    size_t fill_;
    bool simd_;
};

// transcript::Transcript (transcript/src/lib.rs:5-35)
class Transcript {
   public:
    void append(const uint8_t* d, size_t n) { h_.update(d, n); }                       // :16-18
    void append_element(const Field& F, const El& e) { uint8_t be[32]; F.to_be32(e, be); h_.update(be, 32); }
    void sample_challenge(uint8_t out[32]) { h_.finalize_reset(out); h_.update(out, 32); }   // :20-25
    El sample_field_element(const Field& F) { uint8_t d[32]; sample_challenge(d); return F.from_be32_mod_order(d); }  // :27-30

   private:
    Keccak256 h_;
};

}  // namespace host
}  // namespace zk
