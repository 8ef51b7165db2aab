"""Generates tests/golden/sop_vectors.json: sum-of-products sumcheck proofs from the pure-Python
big-int oracle (oracle/zkoracle.py: the reference's prover loop, sumcheck/src/prover.rs:33-73, over SumOfProductsPoly).
The reference has no sum of products, so these are restatement outputs, not reference outputs; the C oracle and the
GPU are both checked against this file.  Usage: python tests/golden/make_sop_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import zkoracle as O  # noqa: E402

GKR = [[0, 2], [0, 3], [1, 2, 3]]
CASES = [
    # field, n_vars, n_tables, terms, degree, seed
    (0, 1, 4, GKR, 3, O.DEFAULT_SEED),
    (0, 2, 4, GKR, 3, O.DEFAULT_SEED),
    (0, 5, 4, GKR, 3, O.DEFAULT_SEED),
    (0, 8, 4, GKR, 3, 12345),
    (1, 6, 4, GKR, 3, O.DEFAULT_SEED),
    (0, 6, 3, [[0], [1, 2]], 2, 99),
    (0, 5, 2, [[0, 0], [1]], 2, 7),
    (1, 5, 5, [[0, 1, 2, 3], [4], [2, 4]], 4, 8),
    (0, 6, 4, GKR, 2, 5),
    (0, 7, 1, [[0]], 1, 6),
    (0, 6, 8, [[0, 1], [2, 3], [4, 5], [6, 7], [0, 7], [1, 6], [2, 5], [3, 4]], 2, 4),
]


def main():
    out = {"note": "restatement outputs (not reference outputs): see the generator's docstring", "cases": []}
    for fid, n, nt, terms, d, seed in CASES:
        F = O.FIELDS[fid]
        sp = O.SumOfProductsPoly([O.MultiLinearPolynomial(F, n, O.gen_table(F, seed, 20 + k, n)) for k in range(nt)], terms)
        claim = sum(sp.prod_reduce()) % F.p
        prover = O.SumcheckProver(d)
        proof, ch = prover.prove_partial(sp, claim)
        fin = [q.evaluations[0] for q in prover.final_poly.polynomials]
        hx = lambda v: "%064x" % v
        out["cases"].append({
            "field": fid, "n_vars": n, "n_tables": nt, "terms": terms, "degree": d, "seed": seed,
            "sum": hx(claim), "round_polys": [hx(x) for r in proof.round_polys for x in r],
            "challenges": [hx(c) for c in ch], "final_evals": [hx(v) for v in fin],
        })
    with open(os.path.join(HERE, "sop_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
        f.write("\n")
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
