#!/usr/bin/env python
"""Sum-of-products sumcheck on one GPU: the GKR layer shape add.Wb + add.Wc + mul.Wb.Wc over four
2^n-entry tables, MAX_VAR_DEGREE 3, prove_partial with the tables resident; the (3,3) ProductPoly proof at the same
size beside it.  Times come from the library itself (zk_ctx_last_prove_ms: host clock around the round loop, CUDA
events around every launch) — no torch import.  One JSON line per polynomial.
usage: python scripts/bench_sop.py [log_n=24] [reps=3] [sop]     ("sop": skip the product proof)"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import zk_b200 as zk

GKR = [[0, 2], [0, 3], [1, 2, 3]]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ctx = zk.Context(0)
    d = 3
    polys = [("sum_of_products_gkr", 4, lambda t: zk.SumOfProductsPoly.new(t, GKR)), ("product_3", 3, lambda t: zk.ProductPoly.new(t))]
    if len(sys.argv) > 3 and sys.argv[3] == "sop":
        polys = polys[:1]
    for name, nt, make in polys:
        tabs = [zk.MultiLinearPolynomial.generate(n, 20 + k, ctx=ctx) for k in range(nt)]
        poly = make(tabs)
        claim = poly.sum()
        runs = []
        for i in range(reps + 1):
            for k in range(nt):
                tabs[k].regenerate(20 + k)
            prover = zk.SumcheckProver(d)
            proof, ch = prover.prove_partial(poly, claim)
            ms = ctx.last_prove_ms()
            runs.append((ms["total_ms"], ms["kernel_ms"], ctx.last_round_ms()[:4]))
        runs = sorted(runs[1:])
        tot, ker, first = runs[len(runs) // 2]
        sub = zk.SumcheckVerifier.verify_partial(proof)
        for k in range(nt):
            tabs[k].regenerate(20 + k)
        ok = sub.challenges == ch and sub.sum == poly.evaluate(ch)
        alg_bytes = 32 * nt * ((1 << n) + 1.5 * ((1 << (n + 1)) - 2))
        print(json.dumps({"poly": name, "log_n": n, "n_tables": nt, "degree": d, "prove_ms": tot, "kernel_ms": ker,
                          "first_round_ms": first, "alg_gbs": alg_bytes / (tot * 1e-3) / 1e9,
                          "first_fused_step_gbs": 48 * nt * (1 << n) / (first[1] * 1e-3) / 1e9 if len(first) > 1 else None,
                          "verified_against_evaluate": bool(ok),
                          "fold_pipe": os.environ.get("ZK_B200_SOP_FOLD_PIPE", "default"),
                          "deferred_reduction": os.environ.get("ZK_B200_SOP_WIDE", "0") == "1"}), flush=True)
        del poly, tabs


if __name__ == "__main__":
    main()
