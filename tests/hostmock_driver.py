#!/usr/bin/env python
"""Runs the REAL host orchestration of the C ABI (zk_b200/csrc/api.cu: round loop and transcript hops, derived S(1)
claims, absorb pipeline, verifier, evaluate/partial_evaluate chains, the multi-GPU NTT's step functions with virtual
ranks, buffer swaps with the NTT plan) on a machine WITHOUT a GPU, against the oracle.

How: tests/test_hostmock_orchestration.py compiles api.cu as plain C++ and links it with tests/cpp/hostmock/ — a
stand-in for the two dozen CUDA runtime calls api.cu makes ("device" memory is host memory, launches run inline) and
for the kernel launchers (the replayable kernels run from their real source; the others are naive models of their
documented contract).  This process is started with ZK_B200_LIB pointing at that library.  TEST INFRASTRUCTURE ONLY:
it checks the caller side of every launch; the kernels themselves are checked by the GPU suite.
Prints one JSON line; exit code 0 = all checks passed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ctypes as C

import numpy as np

os.environ.setdefault("ZK_B200_H2D_OVERLAP_MIN_LOG", "5")  # read once by the library: small tables take the sliced upload too

import cref
import zk_b200 as zk
import zkoracle as O
from zk_b200 import _ffi

GKR = [[0, 2], [0, 3], [1, 2, 3]]


def main():
    lib = _ffi.lib()
    assert "hostmock" in _ffi.SO_PATH and C.CDLL(_ffi.SO_PATH).zk_b200_is_host_mock() == 1, "not the host mock"
    ctx = zk.Context(0)
    checks, fails = 0, []

    def check(cond, what):
        nonlocal checks
        checks += 1
        if not cond:
            fails.append(what)

    # ---- generator, conversions, MLE utilities
    for fid in (0, 1):
        t = zk.MultiLinearPolynomial.generate(7, 2, seed=77, field=fid, ctx=ctx)
        ref = cref.gen_table(fid, 77, 2, 7)
        check((t.evaluation_slice_mont() == ref).all(), ("generate", fid))
        check(t.to_bytes() == cref.to_bytes(fid, ref), ("to_bytes", fid))
        FF = O.FIELDS[fid]
        for iv, cnt in [(0, 1), (0, 3), (3, 2), (6, 1), (0, 7)]:
            assigns = [O.gen_element(9, 9, i) % FF.p for i in range(cnt)]
            out = t.partial_evaluate(iv, assigns)
            exp = cref.partial_evaluate(fid, ref, 7, iv, cref.ints_to_mont(fid, assigns))
            check((out.evaluation_slice_mont() == exp).all(), ("partial_evaluate", fid, iv, cnt))
        pt = [O.gen_element(1, 1, i) % FF.p for i in range(7)]
        om = O.MultiLinearPolynomial(FF, 7, cref.mont_to_ints(fid, ref))
        check(t.evaluate(pt) == om.evaluate(pt), ("evaluate", fid))

    # ---- ProductPoly prover / verifier: the round loop of api.cu (prove_core) against the C oracle
    for fid, n, m, d in [(0, 6, 3, 3), (0, 5, 2, 2), (1, 5, 1, 1), (0, 4, 3, 2), (0, 4, 2, 4), (1, 3, 3, 3), (0, 1, 3, 3), (0, 2, 2, 2),
                         (0, 0, 2, 2), (0, 5, 4, 7)]:
        refs = [cref.gen_table(fid, 5, k, n) for k in range(m)]
        rsum = cref.product_sum(fid, refs, n)
        for absorb in (False, True):
            rp, ch, fin = cref.prove(fid, refs, n, d, rsum, absorb)
            tabs = [zk.MultiLinearPolynomial.new(n, r, field=fid, ctx=ctx) for r in refs]
            pp = zk.ProductPoly.new(tabs)
            check((pp.sum_mont() == rsum).all(), ("product_sum", fid, n, m))
            keep = pp.clone()
            prover = zk.SumcheckProver(d)
            claim = zk.from_mont(fid, rsum)[0]
            proof, gch = (prover.prove(pp, claim), None) if absorb else prover.prove_partial(pp, claim)
            check((proof._round_polys_mont == rp).all(), ("prove round polys", fid, n, m, d, absorb))
            if not absorb:
                check(gch == cref.mont_to_ints(fid, ch), ("challenges", fid, n, m, d))
                check(prover.final_evals == cref.mont_to_ints(fid, fin), ("final evals", fid, n, m, d))
            elif d >= m:
                check(zk.SumcheckVerifier.verify(keep, proof) is True, ("verify accepts", fid, n, m, d))
        # a wrong claim only enters the transcript (derived S(1) must not use it)
        if n >= 2:
            wrong = cref.ints_to_mont(fid, [4242])[0]
            rp2, _, _ = cref.prove(fid, refs, n, d, wrong, False)
            tabs = [zk.MultiLinearPolynomial.new(n, r, field=fid, ctx=ctx) for r in refs]
            p2, _ = zk.SumcheckProver(d).prove_partial(zk.ProductPoly.new(tabs), 4242)
            check((p2._round_polys_mont == rp2).all(), ("wrong claim", fid, n, m, d))

    # the step API
    refs = [cref.gen_table(0, 8, k, 6) for k in range(3)]
    tabs = [zk.MultiLinearPolynomial.new(6, r, ctx=ctx) for r in refs]
    pp = zk.ProductPoly.new(tabs)
    rp, ch, _ = cref.prove(0, refs, 6, 3, cref.product_sum(0, refs, 6), False)
    check(pp.round_poly(3) == cref.mont_to_ints(0, rp[0]), "round_poly step")
    r0 = cref.mont_to_ints(0, ch[:1])[0]
    check(pp.fold_then_round_poly(r0, 3) == cref.mont_to_ints(0, rp[1]), "fold_then_round_poly step")

    # host-table entry point
    n, m, d = 6, 3, 3
    refs = [cref.gen_table(0, 3, k, n) for k in range(m)]
    rsum = cref.product_sum(0, refs, n)
    rp, ch, fin = cref.prove(0, refs, n, d, rsum, False)
    ptrs = (C.c_void_p * m)(*[r.ctypes.data for r in refs])
    grp = np.zeros((n, d + 1, 4), dtype=np.uint64); gch = np.zeros((n, 4), dtype=np.uint64); gfin = np.zeros((m, 4), dtype=np.uint64)
    gsum = np.zeros(4, dtype=np.uint64)
    st = lib.zk_sumcheck_prove_host(ctx.h, 0, ptrs, m, n, d, None, 0, grp.ctypes.data, gch.ctypes.data, gfin.ctypes.data, gsum.ctypes.data)
    check(st == 0 and (grp == rp).all() and (gch == ch).all() and (gfin == fin).all() and (gsum == rsum).all(), "prove_host")
    # the sliced upload with round 0 overlapped (ZK_B200_H2D_OVERLAP_MIN_LOG=5 below makes 2^6 .. 2^9 entries take it):
    # a given claim, a wrong claim (it only enters the transcript), degree below / above the factor count, one table
    for (n, m, d, wrong) in [(6, 3, 3, False), (9, 3, 3, True), (7, 2, 3, False), (8, 3, 2, False), (6, 1, 1, False), (5, 2, 2, False)]:
        refs = [cref.gen_table(0, 11 + n, k, n) for k in range(m)]
        rsum = cref.product_sum(0, refs, n)
        if wrong:
            rsum = cref.ints_to_mont(0, [cref.mont_to_ints(0, rsum.reshape(1, 4))[0] + 7])[0]
        rp, ch, fin = cref.prove(0, refs, n, d, rsum, False)
        ptrs = (C.c_void_p * m)(*[r.ctypes.data for r in refs])
        grp = np.zeros((n, d + 1, 4), dtype=np.uint64); gch = np.zeros((n, 4), dtype=np.uint64); gfin = np.zeros((m, 4), dtype=np.uint64)
        st = lib.zk_sumcheck_prove_host(ctx.h, 0, ptrs, m, n, d, rsum.ctypes.data, 0, grp.ctypes.data, gch.ctypes.data, gfin.ctypes.data, None)
        check(st == 0 and (grp == rp).all() and (gch == ch).all() and (gfin == fin).all(), f"prove_host sliced n={n} m={m} d={d}")

    # ---- sum of products: api.cu + the real kernel source
    with open(os.path.join(ROOT, "tests", "golden", "sop_vectors.json")) as f:
        cases = json.load(f)["cases"]
    for case in cases:
        fid, n, nt, terms, d = case["field"], case["n_vars"], case["n_tables"], case["terms"], case["degree"]
        if n > 6:
            continue
        sp = zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.generate(n, 20 + k, seed=case["seed"], field=fid, ctx=ctx) for k in range(nt)], terms)
        claim = sp.sum()
        check("%064x" % claim == case["sum"], ("sop sum", n, nt))
        keep = sp.clone()
        prover = zk.SumcheckProver(d)
        proof, ch = prover.prove_partial(sp, claim)
        check(["%064x" % x for r in proof.round_polys for x in r] == case["round_polys"], ("sop proof", fid, n, nt, d))
        check(["%064x" % v for v in prover.final_evals] == case["final_evals"], ("sop finals", fid, n, nt, d))
        if d >= max(len(t) for t in terms):
            sub = zk.SumcheckVerifier.verify_partial(proof)
            check(sub.sum == keep.evaluate(ch) == keep.combine(prover.final_evals), ("sop subclaim", fid, n, nt, d))
    refs = [cref.gen_table(0, 3, 20 + k, 5) for k in range(4)]
    rsum = cref.sop_sum(0, refs, GKR, 5)
    rp, _, _ = cref.prove_sop(0, refs, GKR, 5, 3, rsum, absorb=True)
    sp = zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.new(5, r, ctx=ctx) for r in refs], GKR)
    keep = sp.clone()
    proof = zk.SumcheckProver(3).prove(sp, zk.from_mont(0, rsum)[0])
    check((proof._round_polys_mont == rp).all(), "sop prove with absorb")
    check(zk.SumcheckVerifier.verify(keep, proof) is True, "sop verify accepts")
    tampered = zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.new(5, r, ctx=ctx) for r in [refs[1], refs[0], refs[2], refs[3]]], GKR)
    try:  # other bytes absorbed -> other challenges: round 0 still passes (it only involves the claimed sum), round 1 cannot
        zk.SumcheckVerifier.verify(tampered, proof)
        check(False, "sop verify: other tables must be rejected")
    except zk.ZkError as e:
        check(e.status == 7, ("sop verify: other tables status", e.status))
    bad_rounds = zk.SumcheckProof.from_values(0, proof.sum, proof.round_polys[:-1])
    try:
        zk.SumcheckVerifier.verify(keep, bad_rounds)
        check(False, "sop verify: round count")
    except zk.ZkError as e:
        check(e.status == 5, ("sop verify: round count status", e.status))
    wrong = zk.SumcheckProof.from_values(0, proof.sum + 1, proof.round_polys)
    try:
        zk.SumcheckVerifier.verify(keep, wrong)
        check(False, "sop verify: wrong sum")
    except zk.ZkError as e:
        check(e.status == 7, ("sop verify: wrong sum status", e.status))

    # random term structures (1..8 terms of 1..8 factors over 1..8 tables, repeated factors, any MAX_VAR_DEGREE 1..4):
    # the library (api.cu + the real kernel source) against the reference-shaped C oracle
    rng = np.random.default_rng(2024)
    for trial in range(40):
        fid = int(rng.integers(0, 2))
        nt = int(rng.integers(1, 9))
        terms = [[int(x) for x in rng.integers(0, nt, size=int(rng.integers(1, 9)))] for _ in range(int(rng.integers(1, 9)))]
        n, d = int(rng.integers(1, 6)), int(rng.integers(1, 5))
        refs = [cref.gen_table(fid, 1000 + trial, 20 + k, n) for k in range(nt)]
        rsum = cref.sop_sum(fid, refs, terms, n)
        rp, ch, fin = cref.prove_sop(fid, refs, terms, n, d, rsum)
        sp = zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.new(n, r, field=fid, ctx=ctx) for r in refs], terms)
        check((sp.sum_mont() == rsum).all(), ("fuzz sop sum", trial, terms))
        prover = zk.SumcheckProver(d)
        proof, gch = prover.prove_partial(sp, zk.from_mont(fid, rsum)[0])
        check((proof._round_polys_mont == rp).all() and gch == cref.mont_to_ints(fid, ch) and prover.final_evals == cref.mont_to_ints(fid, fin),
              ("fuzz sop prove", trial, fid, n, d, terms))

    # ---- NTT: single-GPU entry (buffer swap with the plan), then the multi-GPU step functions with virtual ranks
    for fid in (0, 1):
        for n in (0, 1, 4, 7, 9):
            a = cref.gen_table(fid, 5, 3, n)
            want = cref.fft(fid, a, n, fast=True)
            t = zk.MultiLinearPolynomial.new(n, a, field=fid, ctx=ctx)
            t.ntt()
            check((t.evaluation_slice_mont() == want).all(), ("zk_ntt", fid, n))
            t.ntt(inverse=True)
            check((t.evaluation_slice_mont() == a).all(), ("zk_ntt round trip", fid, n))
        for G in (2, 4, 8):
            g = G.bit_length() - 1
            for n in sorted({2 * g, 2 * g + 1, 8, 10}):
                if n < 2 * g:
                    continue
                a = cref.gen_table(fid, 6, 4, n)
                want = cref.fft(fid, a, n, fast=True)
                t = zk.MultiLinearPolynomial.new(n, a, field=fid, ctx=ctx)
                t.ntt_virtual_sharded(G)
                check((t.evaluation_slice_mont() == want).all(), ("virtual sharded fft", fid, G, n))
                t.ntt_virtual_sharded(G, inverse=True)
                check((t.evaluation_slice_mont() == a).all(), ("virtual sharded round trip", fid, G, n))
                u = zk.MultiLinearPolynomial.new(n, want, field=fid, ctx=ctx)
                u.ntt_virtual_sharded(G, inverse=True)
                check((u.evaluation_slice_mont() == a).all(), ("virtual sharded ifft", fid, G, n))
        # too few points for the rank count, and an unsupported rank count
        t = zk.MultiLinearPolynomial.new(3, cref.gen_table(fid, 1, 1, 3), field=fid, ctx=ctx)
        for G, want_status in ((4, 13), (3, 13)):
            try:
                t.ntt_virtual_sharded(G)
                check(False, ("expected an error", G))
            except zk.ZkError as e:
                check(e.status == want_status, ("error status", G, e.status))
    loc = cref.gen_table(0, 2, 2, 5)
    check((zk.MultiLinearPolynomial.new_local(5, loc, ctx=ctx).evaluation_slice_mont() == loc).all(), "upload_local")

    print(json.dumps({"hostmock_orchestration_ok": not fails, "checks": checks, "failures": [str(f) for f in fails[:10]]}))
    return 0 if not fails else 1


if __name__ == "__main__":
    sys.exit(main())
