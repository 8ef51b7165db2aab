#!/usr/bin/env python
"""NTT / INTT sweep (BASELINE config 5): 2^16 .. 2^28 points, 1 GPU, tables resident; writes JSON lines.
Timed: zk_ntt on a device-resident table (plan = twiddle table cached), CUDA events on the library stream."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import zk_b200 as zk
from zk_b200 import _ffi

lib = _ffi.lib()
ctx = zk.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream_ptr())
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 28
for field in (0, 1):
    for k in range(lo, hi + 1, 2):
        t = zk.MultiLinearPolynomial.generate(k, 0, seed=3, field=field, ctx=ctx)
        res = {}
        for inverse in (0, 1):
            ctx.check(lib.zk_ntt(ctx.h, t._h, inverse))  # warm-up: builds the plan
            times = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ext); ctx.check(lib.zk_ntt(ctx.h, t._h, inverse)); e1.record(ext); e1.synchronize()
                times.append(e0.elapsed_time(e1))
            res["intt_ms" if inverse else "ntt_ms"] = min(times)
        n = 1 << k
        muls = (n // 2) * k
        print(json.dumps({"field": ["bls12_381_fr", "bls12_377_fr"][field], "log_n": k, **res,
                          "butterfly_mul_per_s": muls / (res["ntt_ms"] * 1e-3), "alg_gbs": 64 * n / (res["ntt_ms"] * 1e-3) / 1e9}), flush=True)
        del t
