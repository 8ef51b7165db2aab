"""GPU suite: radix-2 NTT/INTT parity (fft crate) — golden vectors, oracle at small sizes, naive-DFT spot
checks and round trips at sizes the oracle does not reach."""
import numpy as np
import pytest

import zkoracle as O
from conftest import hx

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


@pytest.mark.parametrize("fid,log_n", [(0, 20), (1, 22)])
def test_fft_large_roundtrip_and_spot_dft(zk, ctx, cref, fid, log_n):
    FF = O.FIELDS[fid]
    n = 1 << log_n
    t = zk.MultiLinearPolynomial.generate(log_n, 1, seed=3, field=fid)
    a = t.evaluation_slice_mont()
    fw = zk.fft(a, field=fid)
    assert (zk.ifft(fw, field=fid) == a).all()
    # X[0] = sum a_j ; X[n/2] = sum (-1)^j a_j ; X[1] by direct evaluation on a 2^12 prefix-sparse input
    vals = cref.mont_to_ints(fid, a[:4096])
    sparse = np.zeros_like(a)
    sparse[:4096] = a[:4096]
    fs = zk.fft(sparse, field=fid)
    w = FF.get_root_of_unity(n)
    for i in (0, 1, n // 2, n - 1, 12345 % n):
        exp = sum(v * pow(w, (i * j) % n, FF.p) for j, v in enumerate(vals)) % FF.p
        assert cref.mont_to_ints(fid, fs[i : i + 1])[0] == exp, i


def test_fft_errors(zk, ctx):
    from zk_b200 import _ffi

    big = zk.MultiLinearPolynomial.generate(2, 0)
    # 2^33 > 2^32 two-adic subgroup of BLS12-381 Fr cannot be allocated here; exercise the check via the helper
    out = np.zeros(4, dtype=np.uint64)
    assert _ffi.lib().zk_field_root_of_unity(0, 1 << 33, out.ctypes.data) == 10
    assert zk.fft([5], field=0) == [5] and zk.ifft([5], field=0) == [5]
    del big
