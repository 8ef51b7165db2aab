#!/usr/bin/env python
"""Multi-GPU NTT sweep (zk_ntt_sharded), one rank per GPU under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
        scripts/bench_ntt_sharded.py [lo=20] [hi=28]

Each rank generates its strided shard on the device, times the forward and the inverse transform (barrier + device
synchronisation on both sides, max over ranks), checks the round trip and — up to 2^24 points — its output block against
the single-GPU transform of the whole table on the same GPU.  Rank 0 prints one JSON line per size.
Written after the round-1 GPU budget was spent: NOT yet run on hardware."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import zk_b200 as zk


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    lo = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    hi = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(zk.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    sctx = zk.Context(local, rank=rank, world=world, nccl_id=idt.cpu().numpy().tobytes())
    uctx = zk.Context(local)

    def timed(fn, reps=5):
        best = None
        for _ in range(reps):
            sctx.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            sctx.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = dt.item() if best is None else min(best, dt.item())
        return best * 1e3

    for field in (0, 1):
        for n in range(lo, hi + 1, 2):
            if (1 << n) < world * world:
                continue
            M = (1 << n) // world
            t = zk.MultiLinearPolynomial.generate(n, 0, seed=3, field=field, ctx=sctx)  # the strided shard of table 0
            shard0 = t.evaluation_slice_mont() if n <= 24 else None
            t.ntt_sharded()  # warm-up: plans, twiddle tables, NCCL channels
            t.ntt_sharded(inverse=True)
            ok = True
            if shard0 is not None:
                ok &= bool((t.evaluation_slice_mont() == shard0).all())  # round trip
                t.ntt_sharded()
                full = zk.MultiLinearPolynomial.generate(n, 0, seed=3, field=field, ctx=uctx)
                full.ntt()
                ok &= bool((t.evaluation_slice_mont() == full.evaluation_slice_mont()[rank * M:(rank + 1) * M]).all())
                t.ntt_sharded(inverse=True)
                del full
            fwd = timed(lambda: t.ntt_sharded())
            t.ntt_sharded(inverse=True)
            # the pair keeps the layout contract (strided -> block -> strided) between timed calls
            inv_pair = timed(lambda: (t.ntt_sharded(), t.ntt_sharded(inverse=True)))
            flag = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                muls = ((1 << n) // 2) * n
                print(json.dumps({"field": ["bls12_381_fr", "bls12_377_fr"][field], "log_n": n, "n_gpus": world, "ntt_ms": fwd,
                                  "ntt_plus_intt_ms": inv_pair, "butterfly_mul_per_s": muls / (fwd * 1e-3),
                                  "bit_exact_vs_single_gpu": bool(flag.item()) if n <= 24 else None}), flush=True)
            del t
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
