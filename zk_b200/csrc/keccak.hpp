// keccak.hpp — host Keccak-256 and the Fiat-Shamir transcript (transcript/src/lib.rs:5-35).
// The reference uses sha3 0.10.8 `Keccak256` (original Keccak padding 0x01, rate 136), through
// `update` and `finalize_reset`.  Hashing stays on the host by design (BASELINE.json north_star).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "host_field.hpp"

namespace zk {
namespace host {

class Keccak256 {
   public:
    Keccak256() { reset(); }
    void reset() { std::memset(s_, 0, sizeof s_); fill_ = 0; }
    void update(const uint8_t* d, size_t n) {
        if (fill_) {
            size_t take = kRate - fill_;
            if (take > n) take = n;
            std::memcpy(buf_ + fill_, d, take);
            fill_ += take; d += take; n -= take;
            if (fill_ == kRate) { absorb(buf_); fill_ = 0; }
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
    void permute() {
        static const uint64_t RC[24] = {
            0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
            0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
            0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
            0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
            0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
            0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
        uint64_t* a = s_;
        for (int r = 0; r < 24; r++) {
            // theta
            uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21],
                     c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22], c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23],
                     c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
            uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1),
                     d4 = c3 ^ rol(c0, 1);
            for (int y = 0; y < 25; y += 5) { a[y] ^= d0; a[y + 1] ^= d1; a[y + 2] ^= d2; a[y + 3] ^= d3; a[y + 4] ^= d4; }
            // rho + pi (lane walk)
            uint64_t cur = a[1], t;
            static const int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
            static const int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
            for (int i = 0; i < 24; i++) { int j = piln[i]; t = a[j]; a[j] = rol(cur, rotc[i]); cur = t; }
            // chi
            for (int y = 0; y < 25; y += 5) {
                uint64_t b0 = a[y], b1 = a[y + 1], b2 = a[y + 2], b3 = a[y + 3], b4 = a[y + 4];
                a[y] = b0 ^ (~b1 & b2); a[y + 1] = b1 ^ (~b2 & b3); a[y + 2] = b2 ^ (~b3 & b4);
                a[y + 3] = b3 ^ (~b4 & b0); a[y + 4] = b4 ^ (~b0 & b1);
            }
            a[0] ^= RC[r];
        }
    }
    uint64_t s_[25];
    uint8_t buf_[kRate];
    size_t fill_;
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
