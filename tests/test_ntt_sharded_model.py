"""CPU suite: the multi-GPU NTT factorisation (zk_ntt_sharded) modelled rank by rank over Python
integers and checked against the oracle's `fft` / `ifft` (fft/src/lib.rs:4-19).

Layout contract: rank q of G holds the STRIDED shard a_q[j] = a[j G + q] (the table sharding of this library); the
forward transform leaves the CONTIGUOUS block X[c M .. (c+1) M) on rank c (M = N / G), the inverse maps blocks back to
strided shards.  Steps (forward): local M-point NTT, twiddle by w_N^(q k'), all-to-all of chunks of C = M / G entries,
a G-point DFT across the ranks' values, all-to-all.  The model below is exactly what api.cu orchestrates; the G-point
DFT is the in-register radix-2 DIF of ntt_sharded_kernels.cuh (its device source is replayed on the host separately).
"""
import pytest

import zkoracle as O


def bitrev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def gdft_dif(F, y, w_pows):
    """Radix-2 DIF on G values, the way the kernel does it in registers: out[c] = sum_q w^(q c) y[q]."""
    G, p = len(y), F.p
    y = list(y)
    h = G // 2
    while h >= 1:
        step = G // (2 * h)
        for b in range(0, G, 2 * h):
            for j in range(h):
                u, v = y[b + j], y[b + j + h]
                y[b + j] = (u + v) % p
                y[b + j + h] = ((u - v) * w_pows[(j * step) % G]) % p
        h //= 2
    bits = G.bit_length() - 1
    return [y[bitrev(c, bits)] for c in range(G)]


def local_ntt(F, vals, inverse):
    return O.ifft(F, vals) if inverse else O.fft(F, vals)


def all_to_all(send):
    """send[q][r] = chunk rank q sends to rank r  ->  recv[r][q]"""
    G = len(send)
    return [[send[q][r] for q in range(G)] for r in range(G)]


def sharded_forward(F, shards):
    G, M, p = len(shards), len(shards[0]), F.p
    N, C = G * M, M // G
    wN, wG = F.get_root_of_unity(N), F.get_root_of_unity(G)
    A = [local_ntt(F, shards[q], False) for q in range(G)]
    A = [[(A[q][k] * pow(wN, q * k, p)) % p for k in range(M)] for q in range(G)]
    B = all_to_all([[A[q][r * C:(r + 1) * C] for r in range(G)] for q in range(G)])  # B[r][q][j]
    w_pows = [pow(wG, i, p) for i in range(G)]
    Y = []
    for r in range(G):
        cols = [gdft_dif(F, [B[r][q][j] for q in range(G)], w_pows) for j in range(C)]  # cols[j][c]
        Y.append([[cols[j][c] for j in range(C)] for c in range(G)])                      # Y[r][c][j]
    R = all_to_all(Y)  # R[c][r][j] = X[c M + r C + j]
    return [[x for r in range(G) for x in R[c][r]] for c in range(G)]


def sharded_inverse(F, blocks):
    G, M, p = len(blocks), len(blocks[0]), F.p
    N, C = G * M, M // G
    wN_inv, wG_inv, g_inv = F.inv(F.get_root_of_unity(N)), F.inv(F.get_root_of_unity(G)), F.inv(G % p)
    B = all_to_all([[blocks[c][r * C:(r + 1) * C] for r in range(G)] for c in range(G)])  # B[r][c][j] = X[c M + r C + j]
    w_pows = [pow(wG_inv, i, p) for i in range(G)]
    Y = []
    for r in range(G):
        cols = [[(v * g_inv) % p for v in gdft_dif(F, [B[r][c][j] for c in range(G)], w_pows)] for j in range(C)]
        Y.append([[cols[j][q] for j in range(C)] for q in range(G)])
    W = all_to_all(Y)  # W[q][r][j], k' = r C + j
    out = []
    for q in range(G):
        w = [x for r in range(G) for x in W[q][r]]
        w = [(w[k] * pow(wN_inv, q * k, p)) % p for k in range(M)]
        out.append(local_ntt(F, w, True))
    return out


@pytest.mark.parametrize("fid", [0, 1])
@pytest.mark.parametrize("G", [2, 4, 8])
def test_sharded_ntt_model_matches_the_oracle(fid, G):
    F = O.FIELDS[fid]
    g = G.bit_length() - 1
    for n in range(2 * g, 2 * g + 3):
        N = 1 << n
        a = O.gen_table(F, 77 + n, 3, n)
        X = O.fft(F, a)
        shards = [a[q::G] for q in range(G)]
        blocks = sharded_forward(F, shards)
        assert [x for b in blocks for x in b] == X, (G, n)
        back = sharded_inverse(F, blocks)
        assert back == shards, (G, n)
        assert O.ifft(F, X) == a


def test_gdft_dif_is_the_dft():
    F = O.BLS12_381_FR
    for G in (2, 4, 8):
        w = F.get_root_of_unity(G)
        y = [O.gen_element(1, 2, i) % F.p for i in range(G)]
        assert gdft_dif(F, y, [pow(w, i, F.p) for i in range(G)]) == O.naive_dft(F, y, w)


def test_sharded_ntt_kernel_sources_replayed_on_the_host(tmp_path):
    """zk_b200/csrc/ntt_sharded_kernels.cuh (power table, inter-rank twiddles, G-point DFT) compiled as plain C++ and
    run thread by thread against direct formulas (tests/cpp/test_ntt_sharded_host.cpp)."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_ntt_sharded_host")
    cmd = ["g++", "-std=c++17", "-O2", "-w", "-I", "/usr/local/cuda/include", "-I", os.path.join(root, "zk_b200", "csrc"),
           os.path.join(root, "tests", "cpp", "test_ntt_sharded_host.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300).stdout
    assert "sharded NTT kernels on the host: 0 mismatches" in out, out
