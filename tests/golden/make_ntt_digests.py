#!/usr/bin/env python
"""Full-size NTT golden digests (BASELINE config 5, fft/src/lib.rs:4-46 at sizes no test can run the oracle at).

The CPU oracle's streamlined transform (oracle/cpu_ref.c `zko_fft_fast`; the CPU suite checks it against the
reference-shaped recursion `zko_fft` and the big-int oracle at small sizes) transforms the seeded table
gen_table(field, seed=3, table_id=1, log_n) once; the Keccak-256 of the natural-order output (Montgomery limbs as they
cross the C ABI) is committed to tests/golden/ntt_digests.json together with a handful of output elements.
tests/test_gpu_ntt.py transforms the same seeded table on the GPU and compares: bit-exact parity of the multi-pass
plans (3 passes at 2^24 / 2^26, the 4-pass plan at 2^28) at the sizes BASELINE quotes.

usage: python tests/golden/make_ntt_digests.py [out.json] [field:log_n ...]      (about 20 minutes of one core)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cref  # noqa: E402

SEED, TABLE_ID = 3, 1
DEFAULT = [(0, 24), (1, 24), (0, 26), (1, 26), (0, 28), (1, 28)]
SPOTS = [0, 1, 2, 12345, (1 << 20) + 7]


def main():
    args = sys.argv[1:]
    out_path = args[0] if args and args[0].endswith(".json") else os.path.join(ROOT, "tests", "golden", "ntt_digests.json")
    cases = [tuple(int(x) for x in a.split(":")) for a in args if ":" in a] or DEFAULT
    res = []
    if os.path.exists(out_path):
        res = json.load(open(out_path))["cases"]
    for fid, k in cases:
        if any(c["field_id"] == fid and c["log_n"] == k for c in res):
            continue
        t0 = time.time()
        a = cref.gen_table(fid, SEED, TABLE_ID, k)
        fw = cref.fft(fid, a, k, fast=True)
        n = 1 << k
        spots = sorted({s % n for s in SPOTS} | {n // 2, n - 1})
        res.append({"field_id": fid, "field": ["bls12_381_fr", "bls12_377_fr"][fid], "seed": SEED, "table_id": TABLE_ID, "log_n": k,
                    "fft_keccak": cref.keccak256(fw.tobytes()).hex(),
                    "spots": {str(s): [hex(int(x)) for x in fw[s]] for s in spots},
                    "oracle_seconds": round(time.time() - t0, 1)})
        print({k2: v for k2, v in res[-1].items() if k2 != "spots"}, flush=True)
        del a, fw
        json.dump({"generator": "tests/golden/make_ntt_digests.py", "oracle": "oracle/cpu_ref.c zko_fft_fast", "cases": res},
                  open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
