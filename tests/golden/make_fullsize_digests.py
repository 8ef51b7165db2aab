#!/usr/bin/env python
"""Full-size golden digests: the CPU oracle (oracle/cpu_ref.c, streamlined in-place prover `zko_prove_fast`, whose
proofs the CPU suite checks against the reference-shaped `zko_prove` and the big-int oracle at small sizes) proves
BASELINE's full-size single-GPU workloads once and the Keccak-256 of (round polynomials || challenges) — the bytes
`bench.py` digests as `proof_keccak` — is committed to tests/golden/fullsize_digests.json.  The GPU suite then proves the
same seeded tables at full size and compares digests (tests/test_gpu_fullsize.py): bit-exact parity at the sizes the
oracle cannot be run at inside a test.  Tables: SURVEY.md 8d generator, seed 0x5EED000000000001, table ids 0..m-1,
claim = the true sum.  Takes about ten minutes and 12 GB on one core.

usage: python tests/golden/make_fullsize_digests.py [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cref  # noqa: E402

SEED = 0x5EED000000000001
FIELD = 0  # BLS12-381 Fr
CASES = [  # (name, n, m, D)
    ("C1 shape: single table 2^20", 20, 1, 1),
    ("C2: product of 2 tables 2^24, degree 2", 24, 2, 2),
    ("(3,3) 2^24", 24, 3, 3),
    ("(1,1) 2^26", 26, 1, 1),
    ("(2,2) 2^26", 26, 2, 2),
    ("C3: degree-3 product of 3 tables 2^26", 26, 3, 3),
]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")
    res = []
    for name, n, m, d in CASES:
        t0 = time.time()
        tabs = [cref.gen_table(FIELD, SEED, k, n) for k in range(m)]
        claim = cref.product_sum(FIELD, tabs, n)
        rp, ch, fin = cref.prove(FIELD, tabs, n, d, claim, False, fast=True)
        digest = cref.keccak256(rp.tobytes() + ch.tobytes()).hex()
        res.append({"name": name, "field": "bls12_381_fr", "seed": hex(SEED), "log_n": n, "m": m, "degree": d,
                    "claim_mont_limbs": [hex(int(x)) for x in claim], "proof_keccak": digest,
                    "finals_keccak": cref.keccak256(fin.tobytes()).hex(), "oracle_seconds": round(time.time() - t0, 1)})
        print(res[-1], flush=True)
        del tabs
    json.dump({"generator": "tests/golden/make_fullsize_digests.py", "oracle": "oracle/cpu_ref.c zko_prove_fast", "cases": res},
              open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
