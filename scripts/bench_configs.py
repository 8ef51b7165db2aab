#!/usr/bin/env python
"""All single-GPU BASELINE configs in one run (writes JSON lines to stdout):
C1 n=20 m=1 D=1 prove+verify, C2 n=24 m=2 D=2 prove (absorb split out) + prove_partial, C3 n=26 m=3 D=3
prove_partial, and the (m,D) shapes at n=24.  Times are medians of 5 after 2 warm-ups, CUDA events on the
library stream around the C-ABI call (tables resident, regenerated in place between runs)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import zk_b200 as zk
from zk_b200 import _ffi

lib = _ffi.lib()
ctx = zk.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream_ptr())
SEED = zk.DEFAULT_SEED


def alg_muls(n, m, d):
    return ((d + 1) * (m - 1) + m) * ((1 << n) - 1)


def timed(fn, reps=5, warm=2, before=None, after=None):
    """median run: returns (ms, after() of that same run)"""
    runs = []
    for i in range(warm + reps):
        if before:
            before()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext); fn(); e1.record(ext); e1.synchronize()
        if i >= warm:
            runs.append((e0.elapsed_time(e1), after() if after else None))
    runs.sort(key=lambda x: x[0])
    return runs[len(runs) // 2]


def run(name, n, m, d, absorb, verify=False):
    tabs = [zk.MultiLinearPolynomial.generate(n, k, seed=SEED, ctx=ctx) for k in range(m)]
    claim = zk.ProductPoly(tabs).sum_mont()
    rp = np.zeros((n, d + 1, 4), dtype=np.uint64); ch = np.zeros((n, 4), dtype=np.uint64); fin = np.zeros((m, 4), dtype=np.uint64)
    arr = zk._table_array(tabs)

    def regen():
        for k in range(m):
            tabs[k].regenerate(k, seed=SEED)

    def prove():
        ctx.check(lib.zk_sumcheck_prove(ctx.h, arr, m, d, claim.ctypes.data, int(absorb), rp.ctypes.data, ch.ctypes.data, fin.ctypes.data))

    ms, pm = timed(prove, before=regen, after=ctx.last_prove_ms)
    out = {"config": name, "log_n": n, "m": m, "D": d, "absorb_initial_poly": absorb, "prove_ms": ms, "absorb_ms": pm["absorb_ms"],
           "round_loop_ms": ms - pm["absorb_ms"], "kernel_ms": pm["kernel_ms"], "field_mul_per_s": alg_muls(n, m, d) / ((ms - pm["absorb_ms"]) * 1e-3),
           "alg_gbs": 32 * m * ((1 << n) + 1.5 * ((1 << (n + 1)) - 2)) / ((ms - pm["absorb_ms"]) * 1e-3) / 1e9}
    if verify:
        regen()
        st = {}

        def ver():
            st["rc"] = lib.zk_sumcheck_verify(ctx.h, arr, m, claim.ctypes.data, rp.ctypes.data, n, d)

        out["verify_ms"] = timed(ver, reps=3, warm=1)[0]
        out["verify_ok"] = st["rc"] == 0
    print(json.dumps(out), flush=True)


def run_evaluate(n):
    """The reference's own criterion bench (polynomial/benches/polynomial_evaluation.rs:85-105):
    MultiLinearPolynomial::evaluate at n = 18..21 variables."""
    t = zk.MultiLinearPolynomial.generate(n, 0, seed=SEED, ctx=ctx)
    pt = zk.to_mont(0, [3 + 7 * i for i in range(n)])
    out = np.zeros(4, dtype=np.uint64)

    def ev():
        ctx.check(lib.zk_mle_evaluate(ctx.h, t._h, pt.ctypes.data, n, out.ctypes.data))

    ms, _ = timed(ev, reps=7, warm=2)
    res = {"config": f"MultiLinearPolynomial::evaluate, {n} vars (reference criterion bench evaluate_{n}_vars)", "log_n": n,
           "evaluate_ms": ms, "field_mul_per_s": ((1 << n) - 1) / (ms * 1e-3)}
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import time

        import cref

        ref = cref.gen_table(0, SEED, 0, n)
        t0 = time.perf_counter()
        exp = cref.partial_evaluate(0, ref, n, 0, pt)
        res["cpu_port_ms"] = (time.perf_counter() - t0) * 1e3
        res["bit_exact_vs_cpu"] = bool((exp[0] == out).all())
    except Exception as e:  # the oracle is only the checker here
        res["cpu_port_ms"] = None
    print(json.dumps(res), flush=True)


for nv in (18, 19, 20, 21):
    run_evaluate(nv)
run("C1 single MLE 2^20, prove+verify", 20, 1, 1, True, verify=True)
run("C1 shape, prove_partial", 20, 1, 1, False)
run("C2 product of 2 MLEs 2^24, prove (with absorb)", 24, 2, 2, True)
run("C2 product of 2 MLEs 2^24, prove_partial", 24, 2, 2, False)
run("(1,1) 2^24 prove_partial", 24, 1, 1, False)
run("(3,3) 2^24 prove_partial", 24, 3, 3, False)
run("(2,2) 2^26 prove_partial", 26, 2, 2, False)
run("(1,1) 2^26 prove_partial", 26, 1, 1, False)
run("C3 degree-3 product 2^26, prove_partial", 26, 3, 3, False)
run("generic path (m=3, D=5) 2^22 prove_partial", 22, 3, 5, False)
