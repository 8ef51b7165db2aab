#!/usr/bin/env python
"""The SHARDED entry points of the C ABI executed without GPUs: G ranks = G threads of this process, each with its own
sharded zk_ctx over the host mock (tests/cpp/hostmock: CUDA runtime stand-in, kernel launchers, and an in-process
NCCL stand-in whose collectives are thread rendezvous).  The real api.cu code runs: strided uploads, the exact
all-reduce of the round evaluations + narrowing, the derived S(1) handed to rank 0 only, the residual all-gather at
several thresholds, the same for the sum-of-products prover, and zk_ntt_sharded's send/recv exchanges.  Every rank's
results are compared with the C oracle.  TEST INFRASTRUCTURE ONLY (see tests/test_hostmock_orchestration.py).
Prints one JSON line; exit code 0 = all checks passed."""
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ctypes as C

import numpy as np

import cref
import zk_b200 as zk
from zk_b200 import _ffi

GKR = [[0, 2], [0, 3], [1, 2, 3]]
SEED = 0x5EED000000000001


def rank_main(rank, world, nccl_id, out):
    fails, checks = [], 0

    def check(cond, what):
        nonlocal checks
        checks += 1
        if not cond:
            fails.append((rank,) + tuple(what if isinstance(what, tuple) else (what,)))

    try:
        ctx = zk.Context(0, rank=rank, world=world, nccl_id=nccl_id)
        lib = _ffi.lib()
        # ---- ProductPoly prover, sharded: strided tables, all-reduced round polynomials, residual gather
        for (n, m, d, thr) in [(8, 3, 3, 4096), (8, 3, 3, 1), (9, 3, 3, 16), (7, 2, 2, 4), (7, 1, 1, 2), (6, 3, 2, 8), (8, 2, 3, 1)]:
            if (1 << n) < 2 * world:
                continue
            ctx.set_gather_threshold(thr)
            refs = [cref.gen_table(0, SEED, k, n) for k in range(m)]
            tabs = [zk.MultiLinearPolynomial.generate(n, k, seed=SEED, ctx=ctx) for k in range(m)]
            check((tabs[0].evaluation_slice_mont() == refs[0][rank::world]).all(), ("sharded generate", n))
            up = zk.MultiLinearPolynomial.new(n, refs[1 % m], ctx=ctx)  # strided 2-D upload of a full host table
            check((up.evaluation_slice_mont() == refs[1 % m][rank::world]).all(), ("sharded upload", n))
            rsum = cref.product_sum(0, refs, n)
            pp = zk.ProductPoly(tabs)
            check((pp.sum_mont() == rsum).all(), ("sharded sum", n, m))
            rp, ch, fin = cref.prove(0, refs, n, d, rsum, False, fast=True)
            check(pp.round_poly(d) == cref.mont_to_ints(0, rp[0]), ("sharded round_poly", n, m, d))
            prover = zk.SumcheckProver(d)
            proof, gch = prover.prove_partial(pp, zk.from_mont(0, rsum)[0])
            check((proof._round_polys_mont == rp).all(), ("sharded prove", n, m, d, thr))
            check(gch == cref.mont_to_ints(0, ch) and prover.final_evals == cref.mont_to_ints(0, fin), ("sharded challenges/finals", n, m, d, thr))
        # ---- sum of products, sharded
        for (n, d, thr) in [(8, 3, 4096), (8, 3, 1), (7, 2, 8), (9, 3, 32)]:
            if (1 << n) < 2 * world:
                continue
            ctx.set_gather_threshold(thr)
            refs = [cref.gen_table(0, SEED, 20 + k, n) for k in range(4)]
            sp = zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.generate(n, 20 + k, seed=SEED, ctx=ctx) for k in range(4)], GKR)
            rsum = cref.sop_sum(0, refs, GKR, n)
            check((sp.sum_mont() == rsum).all(), ("sharded sop sum", n))
            rp, ch, fin = cref.prove_sop(0, refs, GKR, n, d, rsum)
            check(sp.round_poly(d) == cref.mont_to_ints(0, rp[0]), ("sharded sop round_poly", n, d))
            prover = zk.SumcheckProver(d)
            proof, gch = prover.prove_partial(sp, zk.from_mont(0, rsum)[0])
            check((proof._round_polys_mont == rp).all(), ("sharded sop prove", n, d, thr))
            check(gch == cref.mont_to_ints(0, ch) and prover.final_evals == cref.mont_to_ints(0, fin), ("sharded sop finals", n, d, thr))
        # ---- multi-GPU NTT: strided shard in, contiguous block out, and back; inverse alone from a block
        g = world.bit_length() - 1
        for fid in (0, 1):
            for n in sorted({2 * g, 2 * g + 1, 8, 10}):
                if n < 2 * g:
                    continue
                full = cref.gen_table(fid, SEED, 3, n)
                want = cref.fft(fid, full, n, fast=True)
                M = (1 << n) // world
                t = zk.MultiLinearPolynomial.new_local(n, full[rank::world], field=fid, ctx=ctx)
                t.ntt_sharded()
                check((t.evaluation_slice_mont() == want[rank * M:(rank + 1) * M]).all(), ("ntt_sharded forward", fid, n))
                t.ntt_sharded(inverse=True)
                check((t.evaluation_slice_mont() == full[rank::world]).all(), ("ntt_sharded round trip", fid, n))
                u = zk.MultiLinearPolynomial.new_local(n, want[rank * M:(rank + 1) * M], field=fid, ctx=ctx)
                u.ntt_sharded(inverse=True)
                check((u.evaluation_slice_mont() == full[rank::world]).all(), ("ntt_sharded inverse", fid, n))
    except Exception as e:  # a failing rank must not leave the others waiting silently: report and let the join time out
        fails.append((rank, "exception", repr(e)))
    out[rank] = (checks, fails)


def main():
    lib = _ffi.lib()
    assert "hostmock" in _ffi.SO_PATH and C.CDLL(_ffi.SO_PATH).zk_b200_is_host_mock() == 1, "not the host mock"
    total, failures = 0, []
    for world in (2, 4, 8):
        nccl_id = zk.nccl_unique_id()
        assert nccl_id[:8] == b"hostmock", "the NCCL that was loaded is not the in-process stand-in"
        out = {}
        threads = [threading.Thread(target=rank_main, args=(r, world, nccl_id, out), daemon=True) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=600)
        if any(t.is_alive() for t in threads) or len(out) != world:
            failures.append((world, "ranks hung or died", sorted(out.keys())))
            break
        for r in range(world):
            total += out[r][0]
            failures += [(world,) + f for f in out[r][1]]
    print(json.dumps({"hostmock_sharded_ok": not failures, "checks": total, "failures": [str(f) for f in failures[:10]]}))
    sys.stdout.flush()
    os._exit(0 if not failures else 1)  # daemon threads of a hung job must not block the exit


if __name__ == "__main__":
    main()
