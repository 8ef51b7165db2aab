"""CPU suite, part 3: the multi-GPU protocol of SURVEY.md 8e / DESIGN.md 7, run with world_size 2 over
`gloo`.  Two processes each hold the strided shard (index = rank mod world) of seeded tables, compute their
partial round polynomials with the oracle's field arithmetic, all-reduce them as 32-bit limbs widened into
64-bit lanes (the exact trick the CUDA path uses with ncclSum), hash redundantly, fold locally, and gather +
re-interleave the residual.  In rounds >= 1 (and D >= m) each rank skips the t = 1 term and publishes
claim_share - S(0) instead, where the share of S_prev(r_prev) is the whole value on rank 0 and zero elsewhere — the
rule the CUDA path uses (api.cu, `derive_s1`); the all-reduce restores S(1).  The transcript must equal the
single-process oracle's.  This checks the host-side
logic of the N > 1 path (shard axis, lane all-reduce exactness, residual interleave); the kernels themselves
are covered by tests/dist_parity.py on real GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, m, d, gather_at, out):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zkoracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F = O.BLS12_381_FR
    p = F.p
    # strided shard of each table: local[j] = global[j*world + rank]
    tabs = [[O.gen_element(O.DEFAULT_SEED, k, j * world + rank) % p for j in range((1 << n) // world)] for k in range(m)]

    def allreduce_elems(vals):
        lanes = []
        for v in vals:
            lanes += [(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
        t = torch.tensor(lanes, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        res = []
        for e in range(len(vals)):
            acc = 0
            for i in range(8):
                acc += int(t[e * 8 + i]) << (32 * i)
            res.append(acc % p)  # carry-propagate + reduce
        return res

    def lagrange_at(ys, x):  # value at x of the polynomial through (t, ys[t]), t = 0..len-1
        acc = 0
        for t, y in enumerate(ys):
            num, den = 1, 1
            for u in range(len(ys)):
                if u != t:
                    num = num * (x - u) % p
                    den = den * (t - u) % p
            acc = (acc + y * num * pow(den, p - 2, p)) % p
        return acc

    def round_poly(tables, claim_share=None):
        half = len(tables[0]) // 2
        S = []
        for t in range(d + 1):
            if t == 1 and claim_share is not None:  # derived: this rank's share of the claim minus its S(0)
                S.append((claim_share - S[0]) % p)
                continue
            acc = 0
            for j in range(half):
                pr = 1
                for T in tables:
                    pr = pr * ((T[j] + t * (T[j + half] - T[j])) % p) % p
                acc = (acc + pr) % p
            S.append(acc)
        return S

    claim = allreduce_elems([sum(np.prod([T[j] for T in tabs], dtype=object) % p for j in range(len(tabs[0]))) % p])[0]
    tr = O.Transcript()
    tr.append(F.to_bytes_be(claim))
    sharded, cur, rps, chs = True, tabs, [], []
    for rnd in range(n):
        if sharded and (len(cur[0]) < 2 or len(cur[0]) <= gather_at):
            # gather the residual shards and re-interleave: global[j*world + q] = local_q[j]
            full = []
            for T in cur:
                gathered = [None] * world
                dist.all_gather_object(gathered, T)
                L = len(T)
                full.append([gathered[g % world][g // world] for g in range(L * world)])
            cur, sharded = full, False
        share = None
        if rnd >= 1 and d >= m and d >= 1:
            prev_at_r = lagrange_at(rps[-1], chs[-1])
            share = prev_at_r if (not sharded or rank == 0) else 0
        S = round_poly(cur, share)
        if sharded:
            S = allreduce_elems(S)
        rps.append(S)
        tr.append(O.field_elements_to_bytes(F, S))
        r = tr.sample_field_element(F)
        chs.append(r)
        half = len(cur[0]) // 2
        cur = [[(T[j] - r * (T[j] - T[j + half])) % p for j in range(half)] for T in cur]
    if rank == 0:
        out["claim"], out["rps"], out["chs"] = claim, rps, chs
    dist.destroy_process_group()


@pytest.mark.parametrize("n,m,d,gather_at", [(6, 3, 3, 4), (5, 2, 2, 1), (4, 1, 1, 8), (3, 3, 2, 1)])
def test_sharded_protocol_matches_single_process_oracle(n, m, d, gather_at):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zkoracle as O

    F = O.BLS12_381_FR
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, m, d, gather_at, out), nprocs=world, join=True)
    pp = O.ProductPoly([O.MultiLinearPolynomial(F, n, O.gen_table(F, O.DEFAULT_SEED, k, n)) for k in range(m)])
    claim = sum(pp.prod_reduce()) % F.p
    proof, chs = O.SumcheckProver(d).prove_partial(pp, claim)
    assert out["claim"] == claim
    assert out["rps"] == proof.round_polys
    assert out["chs"] == chs
