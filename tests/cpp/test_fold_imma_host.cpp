// Host model of the tensor-path fold (zk_b200/csrc/fold_imma.cuh): the REAL table builder (fixed_mul_table_i8, fragment
// order) and a CPU restatement of the data flow of fe_fold_imma_n for one warp of 32 items — item-major staging rows with the
// chunk swizzle, ldmatrix.x4, mma.m16n8k32.u8.u8.s32 (fragment layouts as the PTX ISA gives them), the pair words, the
// swizzled read-back, the nine-limb assembly, ONE Montgomery row, the conditional subtract and the addition of l — against
// the reference's `left - a * (left - right)` (polynomial/src/multilinear/evaluation_form.rs:68) computed with the
// word-serial host field.  Pins the table layout, the bounds (column sums < 2^21, pair words < 2^30, sum < 2^13 p) and the
// arithmetic in the CPU suite; the device code itself is compared bit for bit on the GPU (tools/imma_fold_probe.cu, the
// parity suite under ZK_B200_SMALL_Q=0).
// Build: g++ -std=c++17 -O2 -I zk_b200/csrc tests/cpp/test_fold_imma_host.cpp -o build/test_fold_imma_host
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fold_imma.cuh"

using namespace zk;

static uint64_t smix(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
    return z;
}
static void limbs(const host::El& e, uint32_t v[8]) {
    for (int k = 0; k < 4; k++) { v[2 * k] = (uint32_t)e.v[k]; v[2 * k + 1] = (uint32_t)(e.v[k] >> 32); }
}

static uint32_t max_col = 0, max_word = 0;

// one warp: out[i] = l[i] + r (h[i] - l[i]) for 32 items through the modelled tensor path
static void warp_fold(const host::Field& F, const host::FieldParams& P, const FixedMulI8& tab, const host::El* l, const host::El* h, host::El* out) {
    alignas(16) unsigned char stage[kImmaStageBytes];
    // every lane stages x = h - l as a 32-byte row, the two 16-byte chunks swapped for rows 4..7 mod 8
    for (int lane = 0; lane < 32; lane++) {
        uint32_t x[8];
        limbs(F.sub(h[lane], l[lane]), x);
        const int sw = (lane >> 2) & 1;
        std::memcpy(stage + lane * 32 + sw * 16, x, 16);
        std::memcpy(stage + lane * 32 + (sw ^ 1) * 16, x + 4, 16);
    }
    // ldmatrix.x4: lane i supplies the address of row (i & 7) of 8 x 16-byte block (i >> 3); lane i receives from block b
    // the 32-bit word (i & 3) of its row (i >> 2)
    uint32_t a[32][2][4];
    for (int T = 0; T < 2; T++) {
        const unsigned char* row_addr[32];
        for (int i = 0; i < 32; i++) {
            const int mi = i >> 3, item = (mi & 1) * 8 + (i & 7), chunk = (mi >> 1) ^ ((item >> 2) & 1);
            row_addr[i] = stage + (16 * T + item) * 32 + chunk * 16;
        }
        for (int i = 0; i < 32; i++)
            for (int b = 0; b < 4; b++) std::memcpy(&a[i][T][b], row_addr[8 * b + (i >> 2)] + 4 * (i & 3), 4);
    }
    // mma.m16n8k32 (A row-major u8, B column-major u8, s32 accumulators), fragment layouts of the PTX ISA
    uint32_t words[32 * 16];
    for (int T = 0; T < 2; T++)
        for (int nt = 0; nt < 4; nt++) {
            unsigned char A[16][32], B[32][8];
            for (int i = 0; i < 32; i++) {
                const int g = i >> 2, t = i & 3;
                for (int b = 0; b < 4; b++) {
                    A[g][4 * t + b] = (unsigned char)(a[i][T][0] >> (8 * b));
                    A[g + 8][4 * t + b] = (unsigned char)(a[i][T][1] >> (8 * b));
                    A[g][16 + 4 * t + b] = (unsigned char)(a[i][T][2] >> (8 * b));
                    A[g + 8][16 + 4 * t + b] = (unsigned char)(a[i][T][3] >> (8 * b));
                    B[4 * t + b][g] = (unsigned char)(tab.frag[i][2 * nt] >> (8 * b));
                    B[16 + 4 * t + b][g] = (unsigned char)(tab.frag[i][2 * nt + 1] >> (8 * b));
                }
            }
            int32_t C[16][8];
            for (int r = 0; r < 16; r++)
                for (int c = 0; c < 8; c++) {
                    int32_t s = 0;
                    for (int k = 0; k < 32; k++) s += (int32_t)A[r][k] * (int32_t)B[k][c];
                    C[r][c] = s;
                    if ((uint32_t)s > max_col) max_col = (uint32_t)s;
                }
            for (int i = 0; i < 32; i++) {  // c0, c1 = row g, columns 2t, 2t+1; c2, c3 = row g + 8
                const int g = i >> 2, t = i & 3;
                const int pc = (nt ^ ((g >> 1) & 3)) * 4 + t;
                words[(16 * T + g) * 16 + pc] = (uint32_t)C[g][2 * t] + ((uint32_t)C[g][2 * t + 1] << 8);
                words[(16 * T + g + 8) * 16 + pc] = (uint32_t)C[g + 8][2 * t] + ((uint32_t)C[g + 8][2 * t + 1] << 8);
            }
        }
    for (int lane = 0; lane < 32; lane++) {
        uint32_t w[16];
        for (int q = 0; q < 4; q++) std::memcpy(w + 4 * q, words + lane * 16 + ((q ^ ((lane >> 1) & 3)) << 2), 16);
        for (int k = 0; k < 16; k++) if (w[k] > max_word) max_word = w[k];
        // nine limbs: even words as they stand + the odd words shifted up by 16 bits
        uint32_t V[10] = {0}, s[9];
        s[0] = w[1] << 16;
        for (int j = 1; j < 8; j++) s[j] = (w[2 * j + 1] << 16) | (w[2 * j - 1] >> 16);
        s[8] = w[15] >> 16;
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)w[2 * j] + s[j]; V[j] = (uint32_t)c; c >>= 32; }
        V[8] = (uint32_t)(c + s[8]);
        // one Montgomery row: m = -V0 (p == 1 mod 2^32)
        const uint32_t m = 0u - V[0];
        c = 0;
        for (int k = 0; k < 8; k++) {
            const uint32_t pk = (uint32_t)(P.p[k / 2] >> (32 * (k & 1)));
            c += (uint64_t)m * pk + V[k];
            V[k] = (uint32_t)c;
            c >>= 32;
        }
        for (int k = 8; k < 10; k++) { c += V[k]; V[k] = (uint32_t)c; c >>= 32; }
        if (V[0] != 0 || V[9] != 0) { std::printf("row did not clear / overflow\n"); std::exit(1); }
        host::El rx;
        for (int k = 0; k < 4; k++) rx.v[k] = (uint64_t)V[1 + 2 * k] | ((uint64_t)V[2 + 2 * k] << 32);
        if (F.geq_p(rx.v)) F.sub_p(rx.v);
        if (F.geq_p(rx.v)) { std::printf("row result >= 2p\n"); std::exit(1); }
        out[lane] = F.add(l[lane], rx);
    }
}

int main() {
    long bad = 0, total = 0;
    for (int fid = 0; fid < 2; fid++) {
        host::Field F(fid);
        const host::FieldParams& P = host::params(fid);
        for (int rr = 0; rr < 8; rr++) {
            host::El r;
            if (rr == 0) r = F.zero();
            else if (rr == 1) r = F.one();
            else if (rr == 2) r = F.neg(F.one());
            else { uint64_t c[4]; for (int k = 0; k < 4; k++) c[k] = smix(91 * rr + k + 1000 * fid); c[3] &= 0x0FFFFFFFFFFFFFFFULL; r = F.from_canonical(c); }
            FixedMulI8 tab;
            fixed_mul_table_i8(F, r, &tab);
            for (int warp = 0; warp < 400; warp++) {
                host::El l[32], h[32], got[32];
                for (int i = 0; i < 32; i++) {
                    for (int k = 0; k < 4; k++) { l[i].v[k] = smix(0x51 + (uint64_t)rr * 7919 + (uint64_t)warp * 256 + i * 8 + k); h[i].v[k] = smix(0x52 + (uint64_t)rr * 7919 + (uint64_t)warp * 256 + i * 8 + 4 + k); }
                    l[i].v[3] &= 0x0FFFFFFFFFFFFFFFULL;
                    h[i].v[3] &= 0x0FFFFFFFFFFFFFFFULL;
                    while (F.geq_p(l[i].v)) F.sub_p(l[i].v);
                    while (F.geq_p(h[i].v)) F.sub_p(h[i].v);
                }
                if (warp == 0) {  // the extremes: x = 0, x = p - 1 (every byte pattern of p - 1), l = p - 1 with x = 1, ...
                    h[0] = l[0];
                    l[1] = F.zero(); h[1] = F.neg(F.one());
                    l[2] = F.neg(F.one()); h[2] = F.zero();
                    l[3] = F.neg(F.one()); h[3] = F.neg(F.one());
                    l[4] = F.zero(); h[4] = F.zero();
                    l[5] = F.one(); h[5] = F.zero();
                }
                warp_fold(F, P, tab, l, h, got);
                for (int i = 0; i < 32; i++) {
                    const host::El want = F.sub(l[i], F.mul(F.sub(l[i], h[i]), r));  // evaluation_form.rs:68
                    total++;
                    if (got[i] != want) { if (bad < 5) std::printf("mismatch field %d r %d warp %d item %d\n", fid, rr, warp, i); bad++; }
                }
            }
        }
    }
    std::printf("max column sum %u (< 2^21 = 2097152), max pair word %u (< 2^30 = 1073741824)\n", max_col, max_word);
    if (max_col >= (1u << 21) || max_word >= (1u << 30)) bad++;
    std::printf("tensor-path fold model: %ld checked, %ld mismatches\n", total, bad);
    return bad != 0;
}
