#!/usr/bin/env python
"""bench.py — sumcheck prove throughput (field-mul/s) and prove time on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU algorithm (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU, NCCL

Workload (config.workload): BASELINE config 3 — degree-3 product sumcheck (3 factor tables, MAX_VAR_DEGREE = 3),
`prove_partial`, BLS12-381 Fr, synthetic seeded tables.  N = 1: 2^26 entries per table (6 GiB, larger than L2).
N > 1: weak scaling, 2^26 entries per GPU (2^(26+log2 N) total), tables sharded by the last-bound variables, the round
sums all-reduced inside the reducing launch (peer mailboxes over NVLink; NCCL all-reduce as the fallback); the same line
carries a `config4` record: BASELINE config 4 exactly (2^30 entries over the N GPUs).  N = 1: an `ntt` record, BASELINE
config 5 (the fft crate's NTT / INTT, 2^16 .. 2^28 points, both fields).  Every proof and transform the line reports is
compared with the CPU oracle's committed digest of the same seeded input (`proof_equals_cpu_oracle_golden`,
`fft_equals_cpu_oracle_golden`).

A step = one whole proof.  `value` = ALG_MULS / prove time with tables resident in HBM (tables are
regenerated on the device between steps, outside the timed region, because prove consumes them);
ALG_MULS(n,m,D) = ((D+1)(m-1)+m)(2^n-1)  (SURVEY.md 8d).  `e2e` = the same metric through
zk_sumcheck_prove_host with HOST (pinned) tables: H2D of every table and D2H of the proof inside the timed region.
`--impl reference`: the reference's CPU algorithm (the reference-shaped C port of oracle/, one core like the reference) on a
bounded sample of the same workload, named in `config.sampled_log_n`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIELD = 0  # BLS12-381 Fr
SEED = 0x5EED000000000001


def alg_muls(n, m, d):
    return ((d + 1) * (m - 1) + m) * ((1 << n) - 1)


def alg_bytes(n, m):
    return 32 * m * ((1 << n) + 1.5 * ((1 << (n + 1)) - 2))


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(n, m, d, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (the first fused step), from the committed
    `ncu --set full` capture of this exact workload (profiles/); None for any other configuration."""
    if not (n - (world.bit_length() - 1) == 26 and m == 3 and d == 3):
        return None, None
    path = os.path.join(ROOT, "profiles", "r02_round_kernels_ncu_full.txt")
    try:
        vals, fused = {}, False
        for line in open(path):
            if line.startswith("Kernel Name"):
                fused = "round_kernel<Fr381, 3, 1," in line  # FOLD = true: the fused fold + round-sum kernel
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if fused and line.startswith(key + " ["):
                    unit = line.split("[")[1].split("]")[0]
                    v = float(line.split("=")[1].strip().replace(",", ""))
                    vals[key] = v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
        return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], os.path.relpath(path, ROOT)
    except Exception:
        return None, None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(n, m, d, steps, warmup):
    """The reference's CPU algorithm (single-threaded, like the reference: no rayon/threads anywhere in it),
    as the reference-shaped C port oracle/cpu_ref.c::zko_prove.  Returns (median s/step, list)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cref

    tabs = [cref.gen_table(FIELD, SEED, k, n) for k in range(m)]
    claim = cref.product_sum(FIELD, tabs, n)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cref.prove(FIELD, tabs, n, d, claim, False, fast=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def cpu_streamlined_run(n, m, d):
    """The streamlined CPU prover (fused single pass per round, in place, pthreads over the pairs) on every host core."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cref

    threads = max(1, min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 64))
    tabs = [cref.gen_table(FIELD, SEED, k, n) for k in range(m)]
    claim = cref.product_sum(FIELD, tabs, n)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        cref.prove(FIELD, tabs, n, d, claim, False, fast=True, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": alg_muls(n, m, d) / best, "unit": "field-mul/s", "cores": threads, "kind": "port-streamlined",
            "sample": f"best of 2 streamlined proofs at 2^{n} entries ({best:.2f} s) on {threads} threads"}


def run_reference(args):
    """The reference arm: the reference's own CPU algorithm (single-threaded like the reference; the reference-shaped C port,
    because the Rust crate cannot be built here) on a BOUNDED sample of the workload.  `config` names the full workload AND
    the sample actually timed (`sampled_log_n`, `extrapolated`): the metric is a rate and the cost is linear in 2^n, so
    the rate of the sample is the rate of the workload; `prove_ms_extrapolated_full` is the full-size time that implies."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    m, d = args.m, args.degree
    n_full = args.log_n if args.log_n else 26 + (args.gpus.bit_length() - 1)
    n = min(args.cpu_log_n, n_full)
    # bounded sample: keep the whole --steps K --warmup W run within a few minutes (~5-11 s per 2^22 proof on one core);
    # every requested warm-up and timed step is run, the sample shrinks instead
    while n > 12 and (args.steps + args.warmup) * 11.0 * (1 << n) / (1 << 22) > 240.0:
        n -= 1
    times = cpu_reference_run(n, m, d, args.steps, args.warmup)
    sec = sum(times) / len(times)
    value = alg_muls(n, m, d) / sec
    sample = f"2^{n}-entry sample of the 2^{n_full}-entry workload (cost is linear in 2^n), {args.warmup} warm-up + {len(times)} timed proofs"
    cfg = workload_config(n_full, m, d, args.gpus)
    cfg.update({"sampled_log_n": n, "extrapolated": n != n_full,
                "sample_note": "the reference arm times a bounded 2^sampled_log_n sample of this workload; rate metrics carry over, times scale by 2^(log_n - sampled_log_n)"})
    line = {
        "impl": "reference", "metric": "sumcheck_prove_field_mul_per_s", "value": value, "unit": "field-mul/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery limbs)",
        "data": "synthetic", "config": cfg, "sampled_log_n": n,
        "cpu_baseline": {"value": value, "unit": "field-mul/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "reference-shaped C restatement (oracle/cpu_ref.c); the Rust reference cannot be built here and is single-threaded"},
        "e2e": {"value": value, "unit": "field-mul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "prove_ms_extrapolated_full": sec * 1e3 * (1 << (n_full - n)),
    }
    if not args.no_ntt:  # config 5 beside it: the reference-shaped recursive fft on one core
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import cref

            a = cref.gen_table(FIELD, 3, 1, 14)
            t0 = time.perf_counter()
            cref.fft(FIELD, a, 14)
            dt = time.perf_counter() - t0
            line["ntt"] = {"metric": "ntt_butterfly_mul_per_s", "value": (1 << 13) * 14 / dt, "unit": "butterfly-mul/s", "cores": 1, "kind": "port",
                           "sample": f"one reference-shaped fft of 2^14 points ({dt:.2f} s; fft/src/lib.rs:21-46 recomputes omega.pow per butterfly)"}
        except Exception as e:
            line["ntt"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n, m, d, gpus):
    return {"workload": f"BASELINE config 3 shape: degree-{d} product sumcheck prove_partial, {m} MLE tables x 2^{n} entries, BLS12-381 Fr"
                        + (" (config 4 sharding)" if gpus > 1 else ""),
            "log_n": n, "n_factors": m, "max_var_degree": d, "field": "bls12_381_fr", "table_bytes_total": 32 * m * (1 << n),
            "sharding": f"strided by last-bound variables over {gpus} GPU(s)", "l2": "inputs larger than L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------------
def ntt_wide_per_transform(k, field):
    """IMAD.WIDE.U32 issued by one 2^k-point transform: one general multiplication per butterfly (112 wide multiplies for
    BLS12-381 Fr, 120 for BLS12-377 Fr, field.cuh) plus one per element and pass boundary (the inter-pass twiddles of the
    four-step factorisation, kernels_ntt.cu)."""
    per_mul = 112 if field == 0 else 120
    passes = (k + 8) // 9
    return per_mul * ((1 << (k - 1)) * k + (passes - 1) * (1 << k))


def ntt_record(args, zk, lib, ctx, ext, torch, np, mb_peak):
    """BASELINE config 5 (fft/src/lib.rs:4-19): zk_ntt forward and inverse on a device-resident seeded table, 2^16 .. 2^28
    points, both fields; CUDA events on the library's stream; L2 flushed between timed transforms of tables that fit it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # larger than the 126 MB L2

    def timed(fn, reps, flush_l2):
        out = []
        for _ in range(reps):
            if flush_l2:
                flush.zero_()
            torch.cuda.synchronize()
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            fn()
            e1.record(ext)
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return out

    golden = {}
    try:
        for c in json.load(open(os.path.join(ROOT, "tests", "golden", "ntt_digests.json")))["cases"]:
            golden[(c["field_id"], c["log_n"])] = c
    except Exception:
        pass
    sweep = []
    for field in (0, 1):
        for k in range(16, args.ntt_max_log_n + 1, 2):
            t = zk.MultiLinearPolynomial.generate(k, 1, seed=3, field=field, ctx=ctx)
            row = {"field": ["bls12_381_fr", "bls12_377_fr"][field], "log_n": k}
            small = (32 << k) <= (128 << 20)
            for inverse in (0, 1):
                # first call builds the plan (the N/2-entry twiddle table): timed once, by the host clock, stated separately
                ctx.synchronize()
                t0 = time.perf_counter()
                ctx.check(lib.zk_ntt(ctx.h, t._h, inverse))
                first_ms = (time.perf_counter() - t0) * 1e3
                timed(lambda: ctx.check(lib.zk_ntt(ctx.h, t._h, inverse)), 3, small)  # warm-up
                ts = timed(lambda: ctx.check(lib.zk_ntt(ctx.h, t._h, inverse)), 5, small)
                key = "intt" if inverse else "ntt"
                row[key + "_ms"] = statistics.median(ts)
                row[key + "_ms_min"] = min(ts)
                row[key + "_first_call_ms_incl_plan_build"] = first_ms
            # after 9 forward + 9 inverse transforms the table is the seeded input again: parity of one more forward pass
            g = golden.get((field, k))
            if g is not None and k <= 26:
                ctx.check(lib.zk_ntt(ctx.h, t._h, 0))
                out = t.evaluation_slice_mont()
                dig = C.create_string_buffer(32)
                lib.zk_keccak256(C.c_void_p(out.ctypes.data), C.c_size_t(out.nbytes), C.cast(dig, C.c_void_p))
                row["fft_equals_cpu_oracle_golden"] = bool(dig.raw.hex() == g["fft_keccak"])
                del out
            muls = (1 << (k - 1)) * k
            sec = row["ntt_ms"] * 1e-3
            row["butterfly_mul_per_s"] = muls / sec
            row["alg_gbs"] = 64.0 * (1 << k) / sec / 1e9  # one read + one write of the table: the algorithmic minimum
            row["imad_wide_per_s"] = ntt_wide_per_transform(k, field) / sec
            row["int_pipe_frac"] = (row["imad_wide_per_s"] / mb_peak) if mb_peak else None
            row["l2"] = "flushed between timed transforms" if small else "table larger than L2"
            sweep.append(row)
            del t
    head = [r for r in sweep if r["field"] == "bls12_381_fr"][-1]
    # end to end through the reference-facing call fft(Vec<F>) -> Vec<F>: host buffer in, host buffer out (zk_ntt_host)
    e2e = None
    ke = min(24, args.ntt_max_log_n)
    p = C.c_void_p()
    if lib.zk_host_alloc(32 << ke, C.byref(p)) == 0:
        src = zk.MultiLinearPolynomial.generate(ke, 1, seed=3, field=FIELD, ctx=ctx)
        ctx.check(lib.zk_table_download(ctx.h, src._h, p))
        del src
        ts = timed(lambda: ctx.check(lib.zk_ntt_host(ctx.h, FIELD, p, 1 << ke, 0)), 2, False)
        ts = timed(lambda: ctx.check(lib.zk_ntt_host(ctx.h, FIELD, p, 1 << ke, 0)), 3, False)
        ms = statistics.median(ts)
        e2e = {"log_n": ke, "ms": ms, "butterfly_mul_per_s": (1 << (ke - 1)) * ke / (ms * 1e-3), "h2d_bytes": 32 << ke, "d2h_bytes": 32 << ke,
               "api": "zk_ntt_host (pinned host vector in, transformed in place, host vector out)"}
        lib.zk_host_free(p)
    cpu = None
    if not args.no_cpu:
        import cref

        a = cref.gen_table(FIELD, 3, 1, 14)
        t0 = time.perf_counter()
        cref.fft(FIELD, a, 14)
        dt = time.perf_counter() - t0
        cpu = {"value": (1 << 13) * 14 / dt, "unit": "butterfly-mul/s", "cores": 1, "kind": "port",
               "sample": f"one reference-shaped fft (recursive, omega.pow per butterfly, fft/src/lib.rs:21-46) of 2^14 points ({dt:.2f} s)"}
        a = cref.gen_table(FIELD, 3, 1, 20)
        t0 = time.perf_counter()
        cref.fft(FIELD, a, 20, fast=True)
        dt = time.perf_counter() - t0
        cpu["streamlined"] = {"value": (1 << 19) * 20 / dt, "unit": "butterfly-mul/s", "cores": 1, "kind": "port-streamlined",
                              "sample": f"iterative radix-2 with a twiddle table, 2^20 points ({dt:.2f} s)"}
    return {"workload": "BASELINE config 5: fft crate radix-2 NTT / INTT (fft/src/lib.rs:4-19), 2^16 .. 2^%d points, natural order in and out, 1 GPU, table resident in HBM" % args.ntt_max_log_n,
            "metric": "ntt_butterfly_mul_per_s", "value": head["butterfly_mul_per_s"], "unit": "butterfly-mul/s", "headline_log_n": head["log_n"],
            "headline_ms": head["ntt_ms"], "timing": "median of 5 after 3 warm-up transforms, CUDA events on the launching stream; the plan (twiddle table) is cached, its one-off build is in *_first_call_ms_incl_plan_build",
            "roofline": {"bound": "int", "kernel": "ntt_pass_kernel (9 radix-2 stages per pass in shared memory)", "achieved": head["imad_wide_per_s"],
                         "peak": mb_peak, "unit": "IMAD.WIDE.U32/s", "frac": head["int_pipe_frac"],
                         "peak_source": "microbench.imad_wide_per_s of this run (independent IMAD.WIDE.U32 chains, zk_b200/csrc/microbench.cu)",
                         "algorithmic_bytes_per_transform": 64 << head["log_n"], "hbm_frac_of_algorithmic_bytes": head["alg_gbs"] / measured_peaks()[0]["hbm_gbs"]},
            "e2e": e2e, "cpu_baseline": cpu, "sweep": sweep}


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    import zk_b200 as zk
    from zk_b200 import _ffi

    lib = _ffi.lib()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(zk.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    ctx = zk.Context(local_rank, rank=rank, world=world, nccl_id=nccl_id)
    m, d = args.m, args.degree
    n = args.log_n if args.log_n else 26 + (world.bit_length() - 1)
    np1 = d + 1
    ext = torch.cuda.ExternalStream(ctx.stream_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def golden_digest(nn):
        """the CPU oracle's digest of the same full-size proof, committed offline (tests/golden/make_fullsize_digests.py)"""
        try:
            for c in json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")))["cases"]:
                if (c["log_n"], c["m"], c["degree"]) == (nn, m, d) and int(c["seed"], 16) == SEED and not c.get("absorb"):
                    return c
        except Exception:
            pass
        return None

    def measure_resident(nn, warmup, steps, sampler=None):
        """`steps` timed proofs of the 2^nn-entry workload with the tables resident in HBM (regenerated between steps,
        untimed: prove consumes them).  CUDA events on the library's stream, barrier + synchronise on both sides."""
        rp = np.zeros((nn, np1, 4), dtype=np.uint64)
        ch = np.zeros((nn, 4), dtype=np.uint64)
        fin = np.zeros((m, 4), dtype=np.uint64)
        tabs = [zk.MultiLinearPolynomial.generate(nn, k, seed=SEED, ctx=ctx) for k in range(m)]
        # the claim (an input of the reference's prove(poly, sum)) — computed once, outside the timed region
        claim = zk.ProductPoly(tabs).sum_mont()
        step_ms, fused_ms, launches0, wall_t0 = [], [], None, None
        for it in range(warmup + steps):
            if it:  # untimed: refill the same allocations
                for k in range(m):
                    tabs[k].regenerate(k, seed=SEED)
            timed = it >= warmup
            if it == 0 and sampler is not None:
                sampler.start()  # started before the warm-up so that nvidia-smi/NVML initialisation is not inside the timed steps
            if timed and launches0 is None:
                launches0 = ctx.launch_count()
                wall_t0 = time.time()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            ctx.check(lib.zk_sumcheck_prove(ctx.h, zk._table_array(tabs), m, d, claim.ctypes.data, 0, rp.ctypes.data, ch.ctypes.data, fin.ctypes.data))
            e1.record(ext)
            barrier()
            if timed:
                step_ms.append(e0.elapsed_time(e1))
                r = ctx.last_round_ms()
                fused_ms.append(r[1] if len(r) > 1 else r[0])
        del tabs
        wall_t1 = time.time()
        launches = ctx.launch_count() - launches0 - (steps - (1 if warmup == 0 else 0)) * m  # minus the (untimed) generator launches
        digest = zk.keccak256(rp.tobytes() + ch.tobytes()).hex()
        g = golden_digest(nn)
        golden_parity = None if g is None else bool(g["proof_keccak"] == digest and g["finals_keccak"] == zk.keccak256(fin.tobytes()).hex()
                                                    and [hex(int(x)) for x in claim] == g["claim_mont_limbs"])
        # the reference verifier's own checks on the last proof (host side, sumcheck/src/verifier.rs:44-78):
        # every round check passes, the replayed challenges equal the prover's, and the subclaim equals the
        # product of the fully folded factors
        sub = np.zeros(4, dtype=np.uint64)
        vch = np.zeros((nn, 4), dtype=np.uint64)
        vst = lib.zk_sumcheck_verify_partial(FIELD, claim.ctypes.data, rp.ctypes.data, nn, d, sub.ctypes.data, vch.ctypes.data)
        prod = zk.to_mont(FIELD, [1])[0].copy()
        for k in range(m):
            lib.zk_field_mul(FIELD, prod.ctypes.data, fin[k].ctypes.data, prod.ctypes.data)
        verified = bool(vst == 0 and (vch == ch).all() and (prod == sub).all())
        t = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps
        return {"ms_per_step": ms, "value": alg_muls(nn, m, d) / (ms * 1e-3), "step_ms": step_ms, "fused_ms": fused_ms, "launches": int(launches),
                "digest": digest, "golden_parity": golden_parity, "golden_oracle": None if g is None else g.get("oracle", "zko_prove_fast"),
                "verified": verified, "round_ms": ctx.last_round_ms(), "wall": (wall_t0, wall_t1), "claim": claim, "rp": rp, "ch": ch, "fin": fin}

    # ---- device-resident metric ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    res = measure_resident(n, args.warmup, args.steps, sampler if rank == 0 else None)
    clocks = sampler.stop(*res["wall"]) if rank == 0 else None
    step_ms, fused_ms, launches = res["step_ms"], res["fused_ms"], res["launches"]
    proof_digest, golden_parity, verified = res["digest"], res["golden_parity"], res["verified"]
    ms_per_step, value = res["ms_per_step"], res["value"]
    claim, rp, ch, fin = res["claim"], res["rp"], res["ch"], res["fin"]
    main_round_ms = res["round_ms"]

    # ---- end-to-end metric: host (pinned) tables in, proof out, through zk_sumcheck_prove_host -----------------
    e2e = None
    if not args.no_e2e:
        local_len = (1 << n) // world
        host_bytes = 32 * local_len  # each rank stages ITS shard (entries rank, rank+world, ...) in pinned host memory
        ptrs = []
        host_cores = None if os.environ.get("ZK_BENCH_NO_BIND") else zk.bind_host_to_gpu(local_rank)  # pinned buffers next to this GPU's PCIe root
        for k in range(m):
            p = C.c_void_p()
            if lib.zk_host_alloc(host_bytes, C.byref(p)) != 0:
                ptrs = None
                break
            ptrs.append(p)
        if ptrs:
            if True:
                for k in range(m):
                    t_loc = zk.MultiLinearPolynomial.generate(n, k, seed=SEED, ctx=ctx)
                    ctx.check(lib.zk_table_download(ctx.h, t_loc._h, ptrs[k]))
                    del t_loc
                arr = (C.c_void_p * m)(*[p.value for p in ptrs])
                e2e_ms = []
                for it in range(min(args.warmup, 2) + args.steps):
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(ext)
                    ctx.check(lib.zk_sumcheck_prove_host(ctx.h, FIELD, arr, m, n, d, claim.ctypes.data, 0, rp.ctypes.data,
                                                         ch.ctypes.data, fin.ctypes.data, None))
                    e1.record(ext)
                    barrier()
                    if it >= min(args.warmup, 2):
                        e2e_ms.append(e0.elapsed_time(e1))
                assert zk.keccak256(rp.tobytes() + ch.tobytes()).hex() == proof_digest, "e2e proof differs from the resident proof"
                tt = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
                if dist is not None:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e2e_ms_step = float(tt.item()) / args.steps
                e2e = {"value": alg_muls(n, m, d) / (e2e_ms_step * 1e-3), "unit": "field-mul/s", "ms_per_step": e2e_ms_step,
                       "h2d_bytes_per_step": 32 * m * local_len * world + 32 * world, "d2h_bytes_per_step": world * (rp.nbytes + ch.nbytes + fin.nbytes),
                       "api": "zk_sumcheck_prove_host (pinned host tables -> proof on host)",
                       "h2d_gbs": (32 * m * local_len * world) / max(e2e_ms_step - ms_per_step, 1e-9) / 1e6,
                       "host_cores_bound": (len(host_cores) if host_cores else None)}
                for p in ptrs:
                    lib.zk_host_free(p)

    # ---- BASELINE config 4 exactly: 3 tables x 2^30 entries sharded over this run's GPUs (N > 1 only) ---------------------
    config4 = None
    if world > 1 and not args.log_n and not args.no_c4 and m == 3 and d == 3:
        c4 = measure_resident(30, 1, max(2, min(args.steps, 3)))
        config4 = {"workload": f"BASELINE config 4: degree-3 product sumcheck prove_partial, 3 MLE tables x 2^30 entries (96 GiB) sharded over {world} GPUs",
                   "log_n": 30, "n_gpus": world, "prove_ms": c4["ms_per_step"], "value": c4["value"], "unit": "field-mul/s", "steps": len(c4["step_ms"]),
                   "warmup": 1, "step_ms": [round(x, 3) for x in c4["step_ms"]], "proof_keccak": c4["digest"], "verified": c4["verified"],
                   "proof_equals_cpu_oracle_golden": c4["golden_parity"], "golden_oracle": c4["golden_oracle"],
                   "round_kernel_ms": [round(x, 4) for x in c4["round_ms"]]}

    # ---- BASELINE config 5: the fft crate's NTT / INTT sweep (N = 1 only) -----------------------------------------------
    nan = float("nan")
    mb = None
    if rank == 0:  # instruction-rate probes: the integer / FP64 roofline denominators, re-measured every run
        mb = ctx.microbench(FIELD) if not args.no_microbench else {"fe_mul_per_s": nan, "imad_wide_per_s": nan, "dfma_per_s": nan, "fe_mul_fixed_per_s": nan}
        # block 0's clock64 window over the whole multi-wave launch: not an SM clock (the sampled nvidia-smi clocks are in `clocks`)
        mb.pop("sm_clock_mhz", None)
        mb["source"] = ("zk_microbench_run (zk_b200/csrc/microbench.cu): independent data-dependent IMAD.WIDE.U32 / DFMA chains on every SM of this GPU, "
                        "in this run; SASS of the probe loops and a reference output are committed under profiles/")
    ntt = None
    if world == 1 and not args.no_ntt:
        try:
            ntt = ntt_record(args, zk, lib, ctx, ext, torch, np, None if args.no_microbench else mb["imad_wide_per_s"])
        except Exception as e:  # the secondary record must never cost the headline line
            ntt = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: the first fused fold+round-sum step (reads N, writes N/2 per factor) --------
    peaks, peak_src = measured_peaks()
    local_n0 = (1 << n) // world
    fused_bytes = 48 * m * local_n0
    fused_s = (sum(fused_ms) / len(fused_ms)) * 1e-3
    achieved = fused_bytes / fused_s / 1e9
    fused_muls = (2 * m + (d + 1) * (m - 1)) * (local_n0 // 4)
    # Instructions actually issued per item of the fused step.  Product multiplications (field.cuh): an intermediate
    # product is 112 IMAD.WIDE.U32, the last product of a term 64 (unreduced); the cubic/3-factor case carries its
    # terms at the Toom points (3 intermediate products instead of 4).  Folds: with two or more factors they run on
    # the FP64 pipe (field_f64.cuh: 128 DFMA + 6 IMAD.WIDE each), a single table folds on the integer pipe (76);
    # the headline shape folds on the INT8 tensor path (below).
    # Rounds >= 1 with D >= m derive S(1) from the previous round polynomial: one unreduced product less per item.
    last_terms = d if d >= m else d + 1
    if m == 1:
        prod_wide = 0
    elif m == 3 and d == 3:
        prod_wide = 3 * 112 + last_terms * 64
    else:
        prod_wide = (m - 2) * (d + 1) * 112 + last_terms * 64
    # The degree-3, three-factor step folds on the INT8 tensor path (fold_imma.cuh: 8 IMMA.16832.U8.U8 per WARP and fold,
    # 6 IMAD.WIDE per thread for the Montgomery row); ZK_B200_FOLD_PIPE selects another pipe.
    tensor_ok = m == 3 and d == 3
    fold_pipe = os.environ.get("ZK_B200_FOLD_PIPE", "imma2" if tensor_ok else ("f64" if m >= 2 else "int"))
    if fold_pipe.startswith("im") and not tensor_ok:
        fold_pipe = "f64" if m >= 2 else "int"
    if fold_pipe.startswith("im"):
        fold_wide, fold_dfma, fold_imma = 6, 0, 8
    elif fold_pipe.startswith("f"):
        fold_wide, fold_dfma, fold_imma = 6, 128, 0
    else:
        fold_wide, fold_dfma, fold_imma = 76, 0, 0
    wide_per_item = 2 * m * fold_wide + prod_wide
    dfma_per_item = 2 * m * fold_dfma
    imma_per_warp_item = 2 * m * fold_imma  # one warp = 32 items
    items = local_n0 // 4
    wide_per_s = wide_per_item * items / fused_s
    dfma_per_s = dfma_per_item * items / fused_s
    traffic, traffic_src = ncu_traffic(n, m, d, world)
    roofline = {"bound": "hbm", "kernel": f"round_kernel<Fr381,{d},FOLD=true> (first fused fold+round-sum step, m={m})", "achieved": achieved,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_src,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": fused_bytes, "launch_ms": fused_s * 1e3,
                "int_pipe": {"bound": "IMAD.WIDE.U32 issue rate (half-rate integer multiply pipe)", "imad_wide_per_item": wide_per_item,
                             "achieved_imad_wide_per_s": wide_per_s, "peak_imad_wide_per_s": mb["imad_wide_per_s"],
                             "frac": wide_per_s / mb["imad_wide_per_s"], "field_mul_per_s": fused_muls / fused_s,
                             "frac_textbook_128_per_mul": fused_muls / fused_s * 128 / mb["imad_wide_per_s"],
                             "standalone_fe_mul_per_s": mb["fe_mul_per_s"]},
                "fp64_pipe": {"bound": "DFMA issue rate (the folds: exact FP64 dot products against multiples of the challenge)",
                              "dfma_per_item": dfma_per_item, "achieved_dfma_per_s": dfma_per_s, "peak_dfma_per_s": mb["dfma_per_s"],
                              "frac": dfma_per_s / mb["dfma_per_s"], "standalone_fe_mul_fixed_per_s": mb["fe_mul_fixed_per_s"]},
                "fold_pipe": fold_pipe,
                "tensor_pipe": {"instruction": "IMMA.16832.U8.U8 (mma.sync.m16n8k32.u8.u8.s32: the bytes of h - l against a 32 x 32 byte table of the challenge's multiples, exact s32 sums)",
                                "imma_per_warp_of_32_items": imma_per_warp_item, "achieved_imma_per_s": imma_per_warp_item * (items / 32) / fused_s,
                                "peak_imma_per_s": 1.40e11, "peak_source": "tools/imma_fold_probe.cu on this pool's B200 (profiles/r02_imma_fold_experiment.txt)"},
                # both pipes run concurrently: the time the two instruction streams would need at their measured peak
                # issue rates, whichever is longer, over the measured launch time
                "pipe_frac": max(wide_per_s / mb["imad_wide_per_s"], dfma_per_s / mb["dfma_per_s"]),
                "whole_prove": {"alg_bytes": alg_bytes(n, m) / world, "gbs": alg_bytes(n, m) / world / (ms_per_step * 1e-3) / 1e9}}
    # the second kernel of the step: round 0 (no fold): reads every table once, multiplies only — bound by the integer-multiply pipe
    if len(main_round_ms) > 1 and m >= 2:
        r0_s = main_round_ms[0] * 1e-3
        r0_items = local_n0 // 2
        if m == 3 and d == 3:
            r0_wide = 3 * 112 + 4 * 64  # Toom point set: 3 reduced products + 4 unreduced last products
        else:
            r0_wide = (m - 2) * (d + 1) * 112 + (d + 1) * 64
        roofline["round0"] = {"kernel": f"round_kernel<Fr381,{d},FOLD=false> (round 0, m={m})", "launch_ms": main_round_ms[0],
                              "bound": "IMAD.WIDE.U32 issue rate", "imad_wide_per_item": r0_wide,
                              "achieved_imad_wide_per_s": r0_wide * r0_items / r0_s, "peak_imad_wide_per_s": mb["imad_wide_per_s"],
                              "frac": r0_wide * r0_items / r0_s / mb["imad_wide_per_s"],
                              "hbm_gbs": 32 * m * local_n0 / r0_s / 1e9, "hbm_frac": 32 * m * local_n0 / r0_s / 1e9 / peaks["hbm_gbs"]}

    # ---- CPU baseline beside it (bounded sample, rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        ncpu = args.cpu_log_n
        times = cpu_reference_run(ncpu, m, d, 1, 0)
        cpu = {"value": alg_muls(ncpu, m, d) / times[0], "unit": "field-mul/s", "cores": 1, "kind": "port",
               "sample": f"one reference-shaped proof at 2^{ncpu} entries ({times[0]:.1f} s); cost is linear in 2^n",
               "host_cores_available": os.cpu_count()}
        # Beside the faithful single-threaded number: the same algorithm streamlined (one fused pass per round, in place)
        # on all host cores — what a tuned multi-core CPU port does; bit-identical proof (oracle/cpu_ref.c::zko_prove_fast_mt).
        try:
            cpu["streamlined_all_cores"] = cpu_streamlined_run(ncpu, m, d)
        except Exception as e:  # never let the extra baseline break the bench line
            cpu["streamlined_all_cores"] = {"error": repr(e)}

    line = {
        "metric": "sumcheck_prove_field_mul_per_s", "value": value, "unit": "field-mul/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u256 (8x32-bit Montgomery limbs; IMAD.WIDE products, folds as exact INT8 tensor-core / FP64 dot products)", "data": "synthetic",
        "config": workload_config(n, m, d, world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "cpu_baseline": cpu, "prove_ms": ms_per_step, "proof_keccak": proof_digest, "verified": verified,
        "proof_equals_cpu_oracle_golden": golden_parity, "golden_oracle": res["golden_oracle"],
        "round_kernel_ms": [round(x, 4) for x in main_round_ms], "step_ms": [round(x, 3) for x in step_ms], "microbench": mb,
        "config4": config4, "ntt": ntt, "uses_mailbox": (ctx.uses_mailbox() if world > 1 else None),
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=0, help="total table size 2^n (default 26 + log2(gpus))")
    ap.add_argument("--m", type=int, default=3)
    ap.add_argument("--degree", type=int, default=3)
    ap.add_argument("--cpu-log-n", type=int, default=22, help="size of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-microbench", action="store_true", help="skip the instruction-rate probes (profiling runs)")
    ap.add_argument("--no-ntt", action="store_true", help="skip the NTT sweep record (BASELINE config 5)")
    ap.add_argument("--no-c4", action="store_true", help="skip the 2^30 record at N > 1 (BASELINE config 4)")
    ap.add_argument("--ntt-max-log-n", type=int, default=28)
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
