#!/usr/bin/env python
"""Full-size golden digests: the CPU oracle (oracle/cpu_ref.c, streamlined in-place prover `zko_prove_fast`, whose
proofs the CPU suite checks against the reference-shaped `zko_prove` and the big-int oracle at small sizes) proves
BASELINE's full-size single-GPU workloads once and the Keccak-256 of (round polynomials || challenges) — the bytes
`bench.py` digests as `proof_keccak` — is committed to tests/golden/fullsize_digests.json.  The GPU suite then proves the
same seeded tables at full size and compares digests (tests/test_gpu_fullsize.py): bit-exact parity at the sizes the
oracle cannot be run at inside a test.  Tables: SURVEY.md 8d generator, seed 0x5EED000000000001, table ids 0..m-1,
claim = the true sum.  Takes about ten minutes and 12 GB on one core.

Multi-GPU sizes (BASELINE config 4 and the weak-scaling bench: 2^27 .. 2^30 entries per table, degree 3): the tables do not
fit the host at full size, so `zko_prove_generated_mt` streams the generator through round 0 and the first fold (no
2^n-entry table is ever held) and runs the remaining rounds on the 2^(n-1)-entry tables with all host cores; the CPU
suite checks that prover against the reference-shaped one at small sizes, and the 2^24 / 2^26 entries of this file
re-derived with it are identical.  2^30 needs 48 GiB and about 25 minutes on 8 cores.

usage: python tests/golden/make_fullsize_digests.py [out.json] [--big 27 28 29 30]   (existing entries are kept)"""
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


def big(out_path, sizes):
    """2^27 .. 2^30: the streaming multi-core oracle prover; appended to the existing file."""
    doc = json.load(open(out_path))
    threads = os.cpu_count() or 1
    for n in sizes:
        m = d = 3
        if any(c["log_n"] == n and c["m"] == m and c["degree"] == d for c in doc["cases"]):
            continue
        t0 = time.time()
        claim, rp, ch, fin = cref.prove_generated(FIELD, SEED, m, n, d, threads)
        doc["cases"].append({"name": f"degree-3 product of 3 tables 2^{n}" + (" (C4)" if n == 30 else " (weak scaling, %d GPUs)" % (1 << (n - 26))),
                             "field": "bls12_381_fr", "seed": hex(SEED), "log_n": n, "m": m, "degree": d,
                             "claim_mont_limbs": [hex(int(x)) for x in claim],
                             "proof_keccak": cref.keccak256(rp.tobytes() + ch.tobytes()).hex(),
                             "finals_keccak": cref.keccak256(fin.tobytes()).hex(), "oracle": "zko_prove_generated_mt",
                             "oracle_seconds": round(time.time() - t0, 1)})
        print(doc["cases"][-1], flush=True)
        doc["oracle"] = "oracle/cpu_ref.c zko_prove_fast (2^20 .. 2^26), zko_prove_generated_mt (2^27 .. 2^30)"
        json.dump(doc, open(out_path, "w"), indent=1)


def absorb_cases(out_path):
    """BASELINE config 2 exactly: `prove` (the tables' to_bytes() absorbed first, prover.rs:16-17) of 2 tables x 2^24."""
    doc = json.load(open(out_path))
    for name, n, m, d in [("C2: prove WITH the initial-poly absorb, 2 tables 2^24, degree 2", 24, 2, 2), ("C1: prove WITH the absorb, single table 2^20", 20, 1, 1)]:
        if any(c.get("absorb") and (c["log_n"], c["m"], c["degree"]) == (n, m, d) for c in doc["cases"]):
            continue
        t0 = time.time()
        tabs = [cref.gen_table(FIELD, SEED, k, n) for k in range(m)]
        claim = cref.product_sum(FIELD, tabs, n)
        rp, ch, fin = cref.prove(FIELD, tabs, n, d, claim, True, fast=True)
        doc["cases"].append({"name": name, "field": "bls12_381_fr", "seed": hex(SEED), "log_n": n, "m": m, "degree": d, "absorb": True,
                             "claim_mont_limbs": [hex(int(x)) for x in claim], "proof_keccak": cref.keccak256(rp.tobytes() + ch.tobytes()).hex(),
                             "finals_keccak": cref.keccak256(fin.tobytes()).hex(), "oracle": "zko_prove_fast (absorb)",
                             "oracle_seconds": round(time.time() - t0, 1)})
        print(doc["cases"][-1], flush=True)
        json.dump(doc, open(out_path, "w"), indent=1)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".json") else os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")
    if "--absorb" in sys.argv:
        return absorb_cases(out_path)
    if "--big" in sys.argv:
        return big(out_path, [int(x) for x in sys.argv[sys.argv.index("--big") + 1:]])
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
