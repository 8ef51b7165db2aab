"""GPU suite: radix-2 NTT/INTT parity (fft crate) — golden vectors, oracle at small sizes, naive-DFT spot
checks and round trips at sizes the oracle does not reach."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import zkoracle as O
from conftest import ROOT, hx

pytestmark = pytest.mark.gpu


def test_fft_golden(zk, ctx, golden):
    for c in golden["fft"]:
        vals = [hx(x) for x in c["in"]]
        fw = zk.fft(vals, field=c["field"])
        assert [hex(x) for x in fw] == c["fft"], c["log_n"]
        assert zk.ifft(fw, field=c["field"]) == vals


@pytest.mark.parametrize("fid", [0, 1])
@pytest.mark.parametrize("log_n", [4, 9, 10, 13, 16])
def test_fft_vs_c_oracle(zk, ctx, cref, fid, log_n):
    a = cref.gen_table(fid, 17, 3, log_n)
    fw = zk.fft(a, field=fid)
    assert (fw == cref.fft(fid, a, log_n, fast=True)).all()
    assert (zk.ifft(fw, field=fid) == a).all()
    if log_n <= 10:  # the reference-shaped recursion (pow per butterfly) is only affordable when small
        assert (fw == cref.fft(fid, a, log_n)).all()


def _spot_dft(zk, cref, fid, log_n, a):
    """X[i] of a prefix-sparse input (2^12 non-zero coefficients) against the defining sum, at positions that exercise every
    pass of the plan (low, high, odd, half, last)."""
    FF = O.FIELDS[fid]
    n = 1 << log_n
    vals = cref.mont_to_ints(fid, a[:4096])
    t = zk.MultiLinearPolynomial.new(log_n, np.concatenate([a[:4096], np.zeros((n - 4096, 4), dtype=np.uint64)]), field=fid)
    t.ntt()
    fs = t.evaluation_slice_mont()
    w = FF.get_root_of_unity(n)
    for i in (0, 1, n // 2, n - 1, 12345 % n, (n // 3) | 1, (1 << (log_n - 9)) + 5):
        exp = sum(v * pow(w, (i * j) % n, FF.p) for j, v in enumerate(vals)) % FF.p
        assert cref.mont_to_ints(fid, fs[i : i + 1])[0] == exp, (fid, log_n, i)
    # and a sparse input whose support is strided over the whole table (non-zero a[j * n/4096]): X[i] = sum a_j w^(i j n/4096)
    t2 = np.zeros((n, 4), dtype=np.uint64)
    t2[:: n >> 12] = a[:4096]
    u = zk.MultiLinearPolynomial.new(log_n, t2, field=fid)
    del t2
    u.ntt()
    fs = u.evaluation_slice_mont()
    w12 = FF.get_root_of_unity(4096)
    for i in (0, 1, 4095, 4096, n - 1, 987654321 % n):
        exp = sum(v * pow(w12, (i * j) % 4096, FF.p) for j, v in enumerate(vals)) % FF.p
        assert cref.mont_to_ints(fid, fs[i : i + 1])[0] == exp, (fid, log_n, i, "strided")


@pytest.mark.parametrize("fid,log_n", [(0, 20), (1, 22), (0, 24), (1, 26), (0, 27), (1, 28)])
def test_fft_large_roundtrip_and_spot_dft(zk, ctx, cref, fid, log_n):
    """BASELINE config 5 sizes (fft/src/lib.rs:4-19): ifft(fft(a)) == a on the device-resident seeded table and direct-DFT spot
    checks; 2^27 is the largest 3-pass plan, 2^28 the 4-pass one."""
    t = zk.MultiLinearPolynomial.generate(log_n, 1, seed=3, field=fid)
    a = t.evaluation_slice_mont()
    t.ntt()
    t.ntt(inverse=True)
    back = t.evaluation_slice_mont()
    del t
    assert (back == a).all()
    del back
    _spot_dft(zk, cref, fid, log_n, a)


with open(os.path.join(ROOT, "tests", "golden", "ntt_digests.json")) as _f:
    NTT_CASES = json.load(_f)["cases"]


@pytest.mark.parametrize("case", NTT_CASES, ids=[f"{c['field']}-2^{c['log_n']}" for c in NTT_CASES])
def test_fft_fullsize_digest(zk, ctx, case):
    """The whole natural-order output at 2^24 / 2^26 / 2^28 points equals the CPU oracle's (tests/golden/ntt_digests.json,
    made offline by tests/golden/make_ntt_digests.py with oracle/cpu_ref.c): Keccak-256 of the output limbs + spot elements."""
    from zk_b200 import _ffi

    fid, k = case["field_id"], case["log_n"]
    t = zk.MultiLinearPolynomial.generate(k, case["table_id"], seed=case["seed"], field=fid)
    t.ntt()
    out = t.evaluation_slice_mont()
    for pos, limbs in case["spots"].items():
        assert [hex(int(x)) for x in out[int(pos)]] == limbs, (fid, k, pos)
    dig = C.create_string_buffer(32)
    _ffi.lib().zk_keccak256(C.c_void_p(out.ctypes.data), C.c_size_t(out.nbytes), C.cast(dig, C.c_void_p))
    assert dig.raw.hex() == case["fft_keccak"], (fid, k)
    # zk_ntt_host (the reference-facing fft(Vec<F>) call: host buffer in, host buffer out) on the same input gives the same
    if k <= 24:
        t.regenerate(case["table_id"], seed=case["seed"])
        assert (zk.fft(t.evaluation_slice_mont(), field=fid) == out).all()


@pytest.mark.parametrize("fid", [0, 1])
@pytest.mark.parametrize("G", [2, 4, 8])
def test_multi_gpu_ntt_with_virtual_ranks_on_one_gpu(zk, ctx, cref, fid, G):
    """zk_ntt_virtual_sharded: every kernel and index map of the multi-GPU NTT with G virtual ranks on this
    GPU, bit-exact against the oracle's fft, against zk_ntt, and back through the inverse."""
    g = G.bit_length() - 1
    for n in sorted({2 * g, 2 * g + 1, 10, 14, 18}):
        if n < 2 * g:
            continue
        a = cref.gen_table(fid, 5, 3, n)
        want = cref.fft(fid, a, n, fast=True)
        t = zk.MultiLinearPolynomial.new(n, a, field=fid)
        t.ntt_virtual_sharded(G)
        got = t.evaluation_slice_mont()
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (fid, G, n, "first differing output", int(bad[0]), "of", int(bad.size))
        u = zk.MultiLinearPolynomial.new(n, a, field=fid)
        u.ntt()
        assert (u.evaluation_slice_mont() == got).all(), (fid, G, n, "differs from zk_ntt")
        t.ntt_virtual_sharded(G, inverse=True)
        assert (t.evaluation_slice_mont() == a).all(), (fid, G, n, "round trip")
        v = zk.MultiLinearPolynomial.new(n, want, field=fid)
        v.ntt_virtual_sharded(G, inverse=True)
        assert (v.evaluation_slice_mont() == a).all(), (fid, G, n, "inverse alone")


def test_fft_errors(zk, ctx):
    from zk_b200 import _ffi

    big = zk.MultiLinearPolynomial.generate(2, 0)
    # 2^33 > 2^32 two-adic subgroup of BLS12-381 Fr cannot be allocated here; exercise the check via the helper
    out = np.zeros(4, dtype=np.uint64)
    assert _ffi.lib().zk_field_root_of_unity(0, 1 << 33, out.ctypes.data) == 10
    assert zk.fft([5], field=0) == [5] and zk.ifft([5], field=0) == [5]
    del big
