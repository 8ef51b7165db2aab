#!/usr/bin/env python
"""Single-GPU check of the multi-GPU NTT path (zk_ntt_virtual_sharded: G virtual ranks on one GPU, device-to-device
copies in place of the NCCL exchanges) against the single-GPU NTT and the CPU oracle, forward and inverse, both fields.
Written in round 1 after the GPU budget was spent: NOT yet run on hardware.  Once it passes, its cases move into
tests/test_gpu_ntt.py.  Prints one JSON line; exit code 0 = all checks passed.
usage: python scripts/check_virtual_ntt.py [max_log_n=20]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import cref
import zk_b200 as zk


def main():
    max_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ctx = zk.Context(0)
    checks, failures, timings = 0, [], {}
    for fid in (0, 1):
        for G in (2, 4, 8):
            g = G.bit_length() - 1
            for n in sorted({2 * g, 2 * g + 1, 10, 14, 17, max_n}):
                if n < 2 * g or n > max_n:
                    continue
                a = cref.gen_table(fid, 5, 3, n)
                want = cref.fft(fid, a, n, fast=True)
                t = zk.MultiLinearPolynomial.new(n, a, field=fid, ctx=ctx)
                t0 = time.time()
                t.ntt_virtual_sharded(G)
                timings[f"f{fid}_G{G}_n{n}"] = round((time.time() - t0) * 1e3, 3)
                got = t.evaluation_slice_mont()
                fwd = bool((got == want).all())
                t.ntt_virtual_sharded(G, inverse=True)
                back = bool((t.evaluation_slice_mont() == a).all())
                # against the single-GPU transform as well
                u = zk.MultiLinearPolynomial.new(n, a, field=fid, ctx=ctx)
                u.ntt()
                same = bool((u.evaluation_slice_mont() == got).all())
                checks += 3
                if not (fwd and back and same):
                    bad = np.nonzero((got != want).any(axis=1))[0]
                    failures.append({"field": fid, "G": G, "n": n, "forward": fwd, "round_trip": back, "equals_zk_ntt": same,
                                     "first_bad_index": int(bad[0]) if bad.size else None, "bad_count": int(bad.size)})
    print(json.dumps({"virtual_sharded_ntt_ok": not failures, "checks": checks, "failures": failures[:8], "ms": timings}))
    return 0 if not failures else 1


if __name__ == "__main__":
    sys.exit(main())
