// sop_group.hpp — host-side common-factor extraction for the sum-of-products kernels (kernels_sop.cu), in its own header so
// that the CPU suite can run it (tests/cpp/test_sop_kernel_host.cpp replays the kernel on grouped specs against a naive
// model of the ORIGINAL terms).  Needs kernels.h (SopSpec).
#pragma once

namespace zk {

// x.a + x.b -> x.(a + b): two terms of the same length (>= 2) whose factor multisets differ in exactly one REAL table
// become one term whose last factor is the virtual table a + b.  Exact (distributivity), one product per evaluation
// point and item less for every pair found.  Greedy, at most kMaxVirtual pairs.
inline SopSpec sop_group(const SopSpec& in) {
    SopSpec s = in;
    bool again = true;
    while (again && s.n_virt < kMaxVirtual) {
        again = false;
        for (int a = 0; a < s.n_terms && !again; a++)
            for (int b = a + 1; b < s.n_terms && !again; b++) {
                const int len = s.len[a];
                if (len < 2 || s.len[b] != len) continue;
                // multiset difference
                int cnt_a[kMaxFactors + kMaxVirtual] = {0}, cnt_b[kMaxFactors + kMaxVirtual] = {0};
                for (int i = 0; i < len; i++) { cnt_a[s.fac[a][i]]++; cnt_b[s.fac[b][i]]++; }
                int only_a = -1, only_b = -1, diff = 0;
                for (int k = 0; k < kMaxFactors + kMaxVirtual; k++) {
                    const int d = cnt_a[k] - cnt_b[k];
                    if (d == 1 && only_a < 0) only_a = k;
                    else if (d == -1 && only_b < 0) only_b = k;
                    else if (d != 0) diff = 99;
                    diff += d > 0 ? d : -d;
                }
                if (diff != 2 || only_a < 0 || only_b < 0 || only_a >= s.n_tables || only_b >= s.n_tables) continue;
                const int v = s.n_virt++;
                s.virt_a[v] = (uint8_t)only_a;
                s.virt_b[v] = (uint8_t)only_b;
                // term a := common factors, then the virtual table; term b disappears
                uint8_t nf[kMaxFactors];
                int n = 0;
                bool dropped = false;
                for (int i = 0; i < len; i++) {
                    if (!dropped && s.fac[a][i] == only_a) { dropped = true; continue; }
                    nf[n++] = s.fac[a][i];
                }
                nf[n++] = (uint8_t)(s.n_tables + v);
                for (int i = 0; i < len; i++) s.fac[a][i] = nf[i];
                for (int t = b; t + 1 < s.n_terms; t++) {
                    s.len[t] = s.len[t + 1];
                    for (int i = 0; i < kMaxFactors; i++) s.fac[t][i] = s.fac[t + 1][i];
                }
                s.n_terms--;
                again = true;
            }
    }
    return s;
}

}  // namespace zk
