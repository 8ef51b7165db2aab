// Differential test of the AVX-512 absorb loop (zk_b200/csrc/keccak_avx512.cpp) against the portable Keccak-f[1600] of
// keccak.hpp: Keccak-256 KATs, every message length 0..1100 in several chunkings, large random messages; prints both
// throughputs.  Exits 0 with "skipped" when the CPU has no AVX-512.
// Build: g++ -std=c++17 -O2 -I zk_b200/csrc tests/cpp/test_keccak_avx512.cpp build/keccak_avx512.o -o build/test_keccak_avx512
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "keccak.hpp"

using zk::host::Keccak256;

static uint64_t smix(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
    return z;
}
static void digest(bool simd, const uint8_t* d, size_t n, size_t chunk, uint8_t out[32]) {
    Keccak256 h(simd);
    if (chunk == 0) h.update(d, n);
    else for (size_t o = 0; o < n; o += chunk) h.update(d + o, (n - o < chunk) ? n - o : chunk);
    h.finalize_reset(out);
}
static void hex(const uint8_t* d, char* s) { for (int i = 0; i < 32; i++) std::sprintf(s + 2 * i, "%02x", d[i]); }

int main() {
    if (!zk_keccak_avx512_available()) { std::printf("keccak avx512: skipped (no AVX-512 on this CPU)\n"); return 0; }
    long bad = 0, total = 0;
    uint8_t a[32], b[32];
    char s[65];
    digest(true, (const uint8_t*)"", 0, 0, a); hex(a, s);
    if (std::strcmp(s, "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470")) { std::printf("KAT empty: %s\n", s); bad++; }
    digest(true, (const uint8_t*)"abc", 3, 0, a); hex(a, s);
    if (std::strcmp(s, "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45")) { std::printf("KAT abc: %s\n", s); bad++; }
    std::vector<uint8_t> buf(1 << 24);
    for (size_t i = 0; i < buf.size(); i += 8) { uint64_t v = smix(0x5EED + i); std::memcpy(&buf[i], &v, 8); }
    for (size_t n = 0; n <= 1100; n++)
        for (size_t chunk : {(size_t)0, (size_t)1, (size_t)7, (size_t)136, (size_t)137, (size_t)300}) {
            digest(true, buf.data() + n, n, chunk, a);
            digest(false, buf.data() + n, n, chunk, b);
            total++;
            if (std::memcmp(a, b, 32)) { if (bad < 5) std::printf("mismatch n=%zu chunk=%zu\n", n, chunk); bad++; }
        }
    for (int t = 0; t < 40; t++) {
        const size_t n = (size_t)(smix(t) % (1 << 22)) + 1, off = (size_t)(smix(t + 99) % 4096);
        digest(true, buf.data() + off, n, t & 1 ? 65536 + t : 0, a);
        digest(false, buf.data() + off, n, 0, b);
        total++;
        if (std::memcmp(a, b, 32)) { if (bad < 5) std::printf("mismatch big n=%zu\n", n); bad++; }
    }
    // a sponge that keeps absorbing after a squeeze (the transcript's digest chaining, transcript/src/lib.rs:20-25)
    {
        Keccak256 x(true), y(false);
        for (int r = 0; r < 50; r++) {
            x.update(buf.data() + 1000 * r, 4000 + 37 * r); y.update(buf.data() + 1000 * r, 4000 + 37 * r);
            x.finalize_reset(a); y.finalize_reset(b);
            x.update(a, 32); y.update(b, 32);
            total++;
            if (std::memcmp(a, b, 32)) { bad++; break; }
        }
    }
    for (int simd = 0; simd < 2; simd++) {
        auto t0 = std::chrono::steady_clock::now();
        digest(simd != 0, buf.data(), buf.size(), 0, a);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("%s: %.3f GB/s\n", simd ? "avx512  " : "portable", buf.size() / sec / 1e9);
    }
    std::printf("keccak avx512: %ld checked, %ld mismatches\n", total, bad);
    return bad != 0;
}
