#define ZK_KARA_HOST
#include "kara_mul256.cuh"
#include <cstdio>
#include <random>
typedef unsigned __int128 u128;
static void ref(uint32_t* out, const uint32_t* a, const uint32_t* b, int n) {
    uint64_t acc[32] = {0};
    for (int i = 0; i < 2 * n; i++) out[i] = 0;
    for (int i = 0; i < n; i++) {
        uint64_t c = 0;
        for (int j = 0; j < n; j++) { uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + c; out[i + j] = (uint32_t)t; c = t >> 32; }
        out[i + n] = (uint32_t)c;
    }
}
int main() {
    std::mt19937_64 g(7);
    long bad = 0, n = 0;
    for (int it = 0; it < 2000000; it++) {
        uint32_t a[8], b[8], o[16], r[16];
        for (int i = 0; i < 8; i++) {
            int m = g() % 8;
            a[i] = m == 0 ? 0 : m == 1 ? 0xffffffffu : (uint32_t)g();
            m = g() % 8;
            b[i] = m == 0 ? 0 : m == 1 ? 0xffffffffu : (uint32_t)g();
        }
        if (it % 5 == 0) for (int i = 0; i < 4; i++) a[i + 4] = a[i];   // zero difference
        if (it % 7 == 0) for (int i = 0; i < 4; i++) b[i] = b[i + 4];
        zk::kara::mul256(o, a, b);
        ref(r, a, b, 8);
        n++;
        for (int i = 0; i < 16; i++) if (o[i] != r[i]) { bad++; break; }
        uint32_t o4[8], r4[8];
        zk::kara::mul128(o4, a, b); ref(r4, a, b, 4);
        for (int i = 0; i < 8; i++) if (o4[i] != r4[i]) { bad++; break; }
    }
    printf("%ld cases, %ld mismatches\n", n, bad);
    return bad != 0;
}
