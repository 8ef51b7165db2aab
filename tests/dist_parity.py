#!/usr/bin/env python
"""Multi-GPU parity driver (run under torchrun, one rank per GPU; NOT collected by pytest).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity.py

Checks, for several (n, m, D) and gather thresholds: the sharded proof (strided sharding by the last-bound
variables + NCCL all-reduce of the round polynomials + residual gather) is bit-identical to the single-GPU
proof and to the CPU oracle; sharded upload/product_sum/round_poly agree too; the same for the sum-of-products prover
(GKR shape); the multi-GPU NTT (zk_ntt_sharded) against the oracle's fft, forward and round trip.  Prints one JSON line.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch
import torch.distributed as dist

import cref
import zk_b200 as zk
from zk_b200 import _ffi


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(zk.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    sctx = zk.Context(local, rank=rank, world=world, nccl_id=idt.cpu().numpy().tobytes())
    uctx = zk.Context(local)  # unsharded context on the same GPU: the 1-GPU proof to compare with
    lib = _ffi.lib()
    seed, checks, ok = 0x5EED000000000001, 0, True

    def prove(ctx, tabs, d, claim):
        n, m = tabs[0].n_vars(), len(tabs)
        rp = np.zeros((n, d + 1, 4), dtype=np.uint64); ch = np.zeros((n, 4), dtype=np.uint64); fin = np.zeros((m, 4), dtype=np.uint64)
        ctx.check(lib.zk_sumcheck_prove(ctx.h, zk._table_array(tabs), m, d, claim.ctypes.data, 0, rp.ctypes.data, ch.ctypes.data, fin.ctypes.data))
        return rp, ch, fin

    for (n, m, d, thr) in [(16, 3, 3, 4096), (16, 3, 3, 1), (16, 3, 3, 1 << 14), (14, 2, 2, 64), (13, 1, 1, 8), (12, 3, 2, 256),
                           (5, 3, 3, 4096), (3, 2, 2, 1), (20, 3, 3, 4096), (22, 2, 2, 4096)]:
        if (1 << n) < world:
            continue
        sctx.set_gather_threshold(thr)
        st = [zk.MultiLinearPolynomial.generate(n, k, seed=seed, ctx=sctx) for k in range(m)]
        ut = [zk.MultiLinearPolynomial.generate(n, k, seed=seed, ctx=uctx) for k in range(m)]
        # sharded generation == strided slice of the full table
        full0 = ut[0].evaluation_slice_mont()
        ok &= bool((st[0].evaluation_slice_mont() == full0[rank::world]).all()); checks += 1
        sclaim = zk.ProductPoly(st).sum_mont(); uclaim = zk.ProductPoly(ut).sum_mont()
        ok &= bool((sclaim == uclaim).all()); checks += 1
        if n <= 16 and (1 << n) // world >= 2:  # one round polynomial through the public step API
            s_rp = np.zeros((d + 1, 4), dtype=np.uint64); u_rp = np.zeros((d + 1, 4), dtype=np.uint64)
            sctx.check(lib.zk_product_round_poly(sctx.h, zk._table_array(st), m, d, s_rp.ctypes.data))
            uctx.check(lib.zk_product_round_poly(uctx.h, zk._table_array(ut), m, d, u_rp.ctypes.data))
            ok &= bool((s_rp == u_rp).all()); checks += 1
            # sharded upload from a full host table
            up = zk.MultiLinearPolynomial.new(n, full0, ctx=sctx)
            ok &= bool((up.evaluation_slice_mont() == full0[rank::world]).all()); checks += 1
        srp, sch, sfin = prove(sctx, st, d, sclaim)
        urp, uch, ufin = prove(uctx, ut, d, uclaim)
        same = bool((srp == urp).all() and (sch == uch).all() and (sfin == ufin).all())
        ok &= same; checks += 1
        if n <= 16:
            refs = [cref.gen_table(0, seed, k, n) for k in range(m)]
            crp, cch, cfin = cref.prove(0, refs, n, d, cref.product_sum(0, refs, n), False, fast=True)
            ok &= bool((srp == crp).all() and (sch == cch).all() and (sfin == cfin).all()); checks += 1
        if rank == 0 and not same:
            print(f"MISMATCH n={n} m={m} d={d} thr={thr}", flush=True)
    # sum of products, GKR layer shape: sharded == single GPU == CPU oracle
    GKR = [[0, 2], [0, 3], [1, 2, 3]]
    for (n, d, thr) in [(16, 3, 4096), (16, 3, 1), (14, 2, 64), (6, 3, 4096), (20, 3, 4096)]:
        if (1 << n) < world:
            continue
        sctx.set_gather_threshold(thr)
        sp = zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.generate(n, 20 + k, seed=seed, ctx=sctx) for k in range(4)], GKR)
        up = zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.generate(n, 20 + k, seed=seed, ctx=uctx) for k in range(4)], GKR)
        sclaim, uclaim = sp.sum_mont(), up.sum_mont()
        ok &= bool((sclaim == uclaim).all()); checks += 1
        if (1 << n) // world >= 2:
            ok &= sp.round_poly(d) == up.round_poly(d); checks += 1
        claim_int = zk.from_mont(0, uclaim)[0]
        sprover, uprover = zk.SumcheckProver(d), zk.SumcheckProver(d)
        sproof, sch = sprover.prove_partial(sp, claim_int)
        uproof, uch = uprover.prove_partial(up, claim_int)
        same = bool((sproof._round_polys_mont == uproof._round_polys_mont).all()) and sch == uch and sprover.final_evals == uprover.final_evals
        ok &= same; checks += 1
        if n <= 16:
            refs = [cref.gen_table(0, seed, 20 + k, n) for k in range(4)]
            crp, cch, cfin = cref.prove_sop(0, refs, GKR, n, d, cref.sop_sum(0, refs, GKR, n))
            ok &= bool((sproof._round_polys_mont == crp).all()) and sprover.final_evals == cref.mont_to_ints(0, cfin); checks += 1
        if rank == 0 and not same:
            print(f"SOP MISMATCH n={n} d={d} thr={thr}", flush=True)
    # multi-GPU NTT (zk_ntt_sharded): strided shard in -> contiguous block of the natural-order transform out, and back
    g = world.bit_length() - 1
    for fid in (0, 1):
        for n in sorted({2 * g, 2 * g + 1, 12, 16, 20}):
            if n < 2 * g or world not in (2, 4, 8):
                continue
            full = cref.gen_table(fid, seed, 3, n)
            want = cref.fft(fid, full, n, fast=True)
            M = (1 << n) // world
            t = zk.MultiLinearPolynomial.new_local(n, full[rank::world], field=fid, ctx=sctx)
            t.ntt_sharded()
            fwd_ok = bool((t.evaluation_slice_mont() == want[rank * M:(rank + 1) * M]).all())
            t.ntt_sharded(inverse=True)
            back_ok = bool((t.evaluation_slice_mont() == full[rank::world]).all())
            ok &= fwd_ok and back_ok; checks += 2
            if not (fwd_ok and back_ok):
                print(f"NTT MISMATCH rank={rank} field={fid} n={n} forward={fwd_ok} round_trip={back_ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"dist_parity": bool(flag.item()), "world": world, "checks": checks, "uses_mailbox": sctx.uses_mailbox()}), flush=True)
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
