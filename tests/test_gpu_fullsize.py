"""GPU suite: bit-exact parity at BASELINE's full single-GPU sizes.

tests/golden/fullsize_digests.json holds, for every single-GPU BASELINE workload (2^20 ... 2^26 entries), the Keccak-256 of
the proof (round polynomials || challenges, Montgomery limbs as they cross the C ABI) and of the final evaluations that
the CPU oracle produced once, offline, from the same seeded tables (tests/golden/make_fullsize_digests.py, ten minutes
of CPU).  Here the CUDA path proves the same tables at full size through the C ABI and must reproduce the digests:
identical round polynomials, challenges and final claims (BASELINE.json north_star) at the sizes the benchmark is
quoted on, not only at the sizes the oracle finishes in seconds."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")) as f:
    CASES = json.load(f)["cases"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_fullsize_proof_digest(zk, ctx, case):
    from zk_b200 import _ffi

    lib = _ffi.lib()
    n, m, d, seed = case["log_n"], case["m"], case["degree"], int(case["seed"], 16)
    if n >= 29:  # the multi-GPU sizes (2^29: weak scaling on 8 GPUs, 2^30: BASELINE config 4) also fit ONE 180 GB B200
        import torch

        free, _ = torch.cuda.mem_get_info(0)
        if free < 32 * m * (1 << n) + (4 << 30):
            pytest.skip("not enough free HBM for the single-GPU run of a multi-GPU size")
    tabs = [zk.MultiLinearPolynomial.generate(n, k, seed=seed) for k in range(m)]
    pp = zk.ProductPoly.new(tabs)
    claim = pp.sum_mont()
    assert [hex(int(x)) for x in claim] == case["claim_mont_limbs"], "claimed sum differs from the oracle's"
    rp = np.zeros((n, d + 1, 4), dtype=np.uint64)
    ch = np.zeros((n, 4), dtype=np.uint64)
    fin = np.zeros((m, 4), dtype=np.uint64)
    absorb = 1 if case.get("absorb") else 0  # `prove`: the tables' to_bytes() go through the host transcript first (prover.rs:16-17)
    ctx.check(lib.zk_sumcheck_prove(ctx.h, pp._arr(), m, d, claim.ctypes.data, absorb, rp.ctypes.data, ch.ctypes.data, fin.ctypes.data))
    assert zk.keccak256(rp.tobytes() + ch.tobytes()).hex() == case["proof_keccak"]
    assert zk.keccak256(fin.tobytes()).hex() == case["finals_keccak"]
    if absorb:  # verify (absorbs the polynomial again, verifier.rs:21-22) accepts it against fresh copies of the tables
        fresh = zk.ProductPoly.new([zk.MultiLinearPolynomial.generate(n, k, seed=seed) for k in range(m)])
        assert lib.zk_sumcheck_verify(ctx.h, fresh._arr(), m, claim.ctypes.data, rp.ctypes.data, n, d) == 0
        return
    # and the proof is one the verifier accepts
    sub = np.zeros(4, dtype=np.uint64)
    vch = np.zeros((n, 4), dtype=np.uint64)
    assert lib.zk_sumcheck_verify_partial(0, claim.ctypes.data, rp.ctypes.data, n, d, sub.ctypes.data, vch.ctypes.data) == 0
    assert (vch == ch).all()
    del tabs, pp
