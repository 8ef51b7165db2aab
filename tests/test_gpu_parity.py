"""GPU suite: bit-exact parity of the CUDA path (through the C ABI) with the oracle on seeded inputs,
with the committed golden vectors, plus size-independent properties at larger sizes and edge cases."""
import numpy as np
import pytest

import zkoracle as O
from conftest import hx

pytestmark = pytest.mark.gpu

F = O.BLS12_381_FR


def gpu_tables(zk, fid, seed, n, m):
    return [zk.MultiLinearPolynomial.generate(n, k, seed=seed, field=fid) for k in range(m)]


# ---- generator + conversions ---------------------------------------------------------------------------
@pytest.mark.parametrize("fid", [0, 1])
def test_generator_matches_oracle(zk, ctx, cref, golden, fid):
    t = zk.MultiLinearPolynomial.generate(10, 2, seed=77, field=fid)
    assert (t.evaluation_slice_mont() == cref.gen_table(fid, 77, 2, 10)).all()
    g = golden["generator"]
    t0 = zk.MultiLinearPolynomial.generate(4, 0)
    assert t0.evaluations[:2] == [hx(g["elem_k0_i0"]), hx(g["elem_k0_i1"])]


@pytest.mark.parametrize("fid", [0, 1])
def test_to_bytes(zk, ctx, cref, fid):
    t = zk.MultiLinearPolynomial.generate(9, 1, seed=5, field=fid)
    assert t.to_bytes() == cref.to_bytes(fid, cref.gen_table(fid, 5, 1, 9))


def test_regenerate_in_place(zk, ctx, cref):
    t = zk.MultiLinearPolynomial.generate(10, 0, seed=1)
    t.regenerate(2, seed=77)
    assert (t.evaluation_slice_mont() == cref.gen_table(0, 77, 2, 10)).all()
    # a table consumed by prove can be refilled and proved again with the same result
    tabs = [zk.MultiLinearPolynomial.generate(11, k, seed=9) for k in range(3)]
    pp = zk.ProductPoly.new(tabs)
    claim = pp.sum()
    p1, c1 = zk.SumcheckProver(3).prove_partial(pp, claim)
    for k in range(3):
        tabs[k].regenerate(k, seed=9)
    p2, c2 = zk.SumcheckProver(3).prove_partial(pp, claim)
    assert p1.round_polys == p2.round_polys and c1 == c2


def test_upload_download_roundtrip(zk, ctx, cref):
    ref = cref.gen_table(0, 123, 0, 12)
    t = zk.MultiLinearPolynomial.new(12, ref)
    assert (t.evaluation_slice_mont() == ref).all()
    edge = [0, 1, F.p - 1, F.p - 2, 2, (F.p - 1) // 2, F.R % F.p, 12345]
    t2 = zk.MultiLinearPolynomial.new(3, edge)
    assert t2.evaluations == edge


# ---- fold / partial_evaluate -----------------------------------------------------------------------------
def test_partial_evaluate_golden(zk, ctx, golden):
    for c in golden["partial_evaluate"]:
        t = zk.MultiLinearPolynomial.generate(c["n"], 0, seed=c["seed"])
        out = t.partial_evaluate(c["initial_var"], c["assignments"])
        assert [hex(x) for x in out.evaluations] == c["out"], c
        assert out.n_vars() == c["n"] - len(c["assignments"])


@pytest.mark.parametrize("fid", [0, 1])
def test_partial_evaluate_vs_oracle_random(zk, ctx, cref, fid):
    rng = np.random.default_rng(1)
    n = 11
    t = zk.MultiLinearPolynomial.generate(n, 0, seed=9, field=fid)
    ref = cref.gen_table(fid, 9, 0, n)
    FF = O.FIELDS[fid]
    for iv, cnt in [(0, 1), (0, 3), (4, 2), (10, 1), (7, 4), (0, 11), (5, 6)]:
        assigns = [int.from_bytes(rng.bytes(32), "little") % FF.p for _ in range(cnt)]
        out = t.partial_evaluate(iv, assigns)
        exp = cref.partial_evaluate(fid, ref, n, iv, cref.ints_to_mont(fid, assigns))
        assert (out.evaluation_slice_mont() == exp).all(), (iv, cnt)
    # edge assignments 0, 1, p-1
    for a in (0, 1, FF.p - 1):
        out = t.partial_evaluate(3, [a])
        assert (out.evaluation_slice_mont() == cref.partial_evaluate(fid, ref, n, 3, cref.ints_to_mont(fid, [a]))).all()


def test_partial_evaluate_range_errors(zk, ctx):
    t = zk.MultiLinearPolynomial.new(2, [1, 2, 3, 4])
    with pytest.raises(zk.ZkError):  # Rust: subtract-with-overflow panic (pairing_index.rs:3,6)
        t.partial_evaluate(1, [1, 2])
    with pytest.raises(zk.ZkError):
        t.partial_evaluate(2, [1])
    assert t.partial_evaluate(0, []).evaluations == [1, 2, 3, 4]
    assert t.partial_evaluate(0, [5, 6]).n_vars() == 0


def test_evaluate_vs_oracle(zk, ctx, cref):
    n = 13
    t = zk.MultiLinearPolynomial.generate(n, 1, seed=4)
    pt = [O.gen_element(8, 0, i) for i in range(n)]
    ref = cref.gen_table(0, 4, 1, n)
    exp = cref.partial_evaluate(0, ref, n, 0, cref.ints_to_mont(0, pt))
    assert t.evaluate(pt) == cref.mont_to_ints(0, exp)[0]
    z = zk.MultiLinearPolynomial.new(0, [42])
    assert z.evaluate([]) == 42


# ---- product ops --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [1, 2, 3, 5])
def test_product_sum_and_prod_reduce(zk, ctx, cref, m):
    n = 12
    tabs = gpu_tables(zk, 0, 21, n, m)
    pp = zk.ProductPoly.new(tabs)
    refs = [cref.gen_table(0, 21, k, n) for k in range(m)]
    assert (pp.sum_mont() == cref.product_sum(0, refs, n)).all()
    pr = cref.mont_to_ints(0, refs[0])
    for k in range(1, m):
        pr = [a * b % F.p for a, b in zip(pr, cref.mont_to_ints(0, refs[k]))]
    assert pp.prod_reduce() == pr


@pytest.mark.parametrize("m,d", [(1, 1), (2, 2), (3, 3), (2, 3), (3, 1), (4, 4), (1, 3), (3, 2)])
def test_round_poly_and_fold_steps_vs_oracle(zk, ctx, cref, m, d):
    """zk_product_round_poly / fold_inplace / fold_then_round_poly each against the oracle's round loop."""
    n = 9
    refs = [cref.gen_table(0, 31, k, n) for k in range(m)]
    claim = cref.product_sum(0, refs, n)
    rp, ch, fin = cref.prove(0, refs, n, d, claim, False, fast=True)
    exp_rp = [cref.mont_to_ints(0, rp[i]) for i in range(n)]
    chs = cref.mont_to_ints(0, ch)
    # unfused: round_poly, fold, round_poly, ...
    pp = zk.ProductPoly.new(gpu_tables(zk, 0, 31, n, m))
    for i in range(n):
        assert pp.round_poly(d) == exp_rp[i], ("unfused", i)
        pp.fold_inplace(chs[i])
        assert pp.n_vars() == n - i - 1
    assert [q.evaluations[0] for q in pp.polynomials] == cref.mont_to_ints(0, fin)
    # fused
    pp = zk.ProductPoly.new(gpu_tables(zk, 0, 31, n, m))
    assert pp.round_poly(d) == exp_rp[0]
    for i in range(1, n):
        assert pp.fold_then_round_poly(chs[i - 1], d) == exp_rp[i], ("fused", i)
    pp.fold_inplace(chs[n - 1])
    assert [q.evaluations[0] for q in pp.polynomials] == cref.mont_to_ints(0, fin)


# ---- full proofs -------------------------------------------------------------------------------------------------
def test_golden_proofs(zk, ctx, golden):
    """Every committed golden proof (reference-test fixtures B.1-B.4 and seeded tables), prove and prove_partial."""
    for c in golden["small_cases"] + golden["seeded_cases"]:
        fid = c["field"]
        if "tables" in c:
            n = len(c["tables"][0]).bit_length() - 1
            tabs = [zk.MultiLinearPolynomial.new(n, t, field=fid) for t in c["tables"]]
        else:
            n = c["n"]
            tabs = gpu_tables(zk, fid, hx(c["seed"]), n, c["m"])
        pp = zk.ProductPoly.new(tabs)
        if "seed" in c:
            assert hex(pp.sum()) == c["claim"], c["name"]
        prover = zk.SumcheckProver(c["degree"])
        if c["absorb"]:
            proof = prover.prove(pp, hx(c["claim"]))
        else:
            proof, ch = prover.prove_partial(pp, hx(c["claim"]))
            assert [hex(x) for x in ch] == c["challenges"], c["name"]
        assert [[hex(x) for x in r] for r in proof.round_polys] == c["round_polys"], c["name"]
        assert [hex(x) for x in prover.final_evals] == c["finals"], c["name"]


@pytest.mark.parametrize("fid,n,m,d", [(0, 16, 1, 1), (0, 16, 2, 2), (0, 16, 3, 3), (1, 14, 3, 3), (1, 13, 2, 2),
                                       (0, 12, 2, 4), (0, 12, 5, 5), (0, 17, 3, 3), (0, 3, 3, 3), (0, 2, 2, 2), (0, 1, 1, 1)])
def test_prove_partial_vs_c_oracle(zk, ctx, cref, fid, n, m, d):
    refs = [cref.gen_table(fid, 0xABC, k, n) for k in range(m)]
    claim = cref.product_sum(fid, refs, n)
    rp, ch, fin = cref.prove(fid, refs, n, d, claim, False, fast=True)
    pp = zk.ProductPoly.new(gpu_tables(zk, fid, 0xABC, n, m))
    assert (pp.sum_mont() == claim).all()
    prover = zk.SumcheckProver(d)
    proof, chs = prover.prove_partial(pp, cref.mont_to_ints(fid, claim.reshape(1, 4))[0])
    assert (proof._round_polys_mont == rp).all()
    assert chs == cref.mont_to_ints(fid, ch)
    assert prover.final_evals == cref.mont_to_ints(fid, fin)


@pytest.mark.parametrize("fid,n,m,d", [(0, 13, 3, 3), (0, 12, 2, 2), (1, 11, 1, 1), (0, 11, 2, 3), (0, 10, 3, 2), (0, 12, 3, 4)])
def test_prove_with_a_wrong_claimed_sum_vs_c_oracle(zk, ctx, cref, fid, n, m, d):
    """The round polynomials are computed from the tables, whatever sum the caller claims (prover.rs:42 only absorbs
    it): a false claim changes the challenges, never the honesty of S_i.  The kernels derive S_i(1) of rounds >= 1 from
    S_{i-1}(r_{i-1}) — an identity of the tables, not of the claim — so a wrong claim (and D != m) must still match."""
    refs = [cref.gen_table(fid, 0xD00D, k, n) for k in range(m)]
    wrong = cref.ints_to_mont(fid, [123456789])[0]
    rp, ch, fin = cref.prove(fid, refs, n, d, wrong, False, fast=False)
    pp = zk.ProductPoly.new(gpu_tables(zk, fid, 0xD00D, n, m))
    prover = zk.SumcheckProver(d)
    proof, chs = prover.prove_partial(pp, 123456789)
    assert (proof._round_polys_mont == rp).all()
    assert chs == cref.mont_to_ints(fid, ch)
    assert prover.final_evals == cref.mont_to_ints(fid, fin)
    if d >= m:
        with pytest.raises(Exception):
            zk.SumcheckVerifier.verify_partial(proof)  # "verifier check failed: claimed_sum != p(0) + p(1)"


def test_zero_variable_and_tiny_products(zk, ctx):
    """n_vars = 0: the prover loop runs zero rounds (`for _ in 0..poly.n_vars()`, prover.rs:44); n = 1: one round."""
    pp = zk.ProductPoly.new([zk.MultiLinearPolynomial.new(0, [5]), zk.MultiLinearPolynomial.new(0, [7])])
    assert pp.sum() == 35 and pp.evaluate([]) == 35
    prover = zk.SumcheckProver(2)
    proof, chs = prover.prove_partial(pp, 35)
    assert proof.round_polys == [] and chs == [] and prover.final_evals == [5, 7]
    sub = zk.SumcheckVerifier.verify_partial(proof)
    assert sub.sum == 35 and sub.challenges == []
    pp1 = zk.ProductPoly.new([zk.MultiLinearPolynomial.new(1, [2, 3]), zk.MultiLinearPolynomial.new(1, [4, 6])])
    O1 = O.ProductPoly([O.MultiLinearPolynomial(F, 1, [2, 3]), O.MultiLinearPolynomial(F, 1, [4, 6])])
    ref, rch = O.SumcheckProver(2).prove_partial(O1, 26)
    proof, chs = zk.SumcheckProver(2).prove_partial(pp1, 26)
    assert proof.round_polys == ref.round_polys and chs == rch


def test_prove_with_absorb_multi_chunk(zk, ctx, cref):
    """prove() on tables larger than one 32 MiB absorb chunk (2^21 entries = 64 MiB per factor): the double-buffered
    device->host serialisation must hash exactly poly.to_bytes()."""
    n = 21
    refs = [cref.gen_table(0, 3, k, n) for k in range(2)]
    claim = cref.product_sum(0, refs, n)
    init = b"".join(cref.to_bytes(0, t) for t in refs)
    rp, ch, fin = cref.prove(0, refs, n, 2, claim, False, fast=True)  # only for shape; transcript differs
    pp = zk.ProductPoly.new(gpu_tables(zk, 0, 3, n, 2))
    claim_int = cref.mont_to_ints(0, claim.reshape(1, 4))[0]
    proof = zk.SumcheckProver(2).prove(pp.clone(), claim_int)
    rc, sub, vch = cref.verify_internal(0, claim, proof._round_polys_mont, init)  # oracle verifier replays the absorb
    assert rc == 0
    assert zk.SumcheckVerifier.verify(pp, proof) is True


def test_config1_prove_verify_2p20(zk, ctx, cref):
    """BASELINE config 1: single random 2^20-entry MLE, D=1, prove + verify, bit-exact vs the CPU oracle."""
    n = 20
    ref = cref.gen_table(0, O.DEFAULT_SEED, 0, n)
    claim = cref.product_sum(0, [ref], n)
    rp, ch, fin = cref.prove(0, [ref], n, 1, claim, True)  # reference-shaped, with the absorb
    t = zk.MultiLinearPolynomial.generate(n, 0)
    pp = zk.ProductPoly.new([t])
    claim_int = pp.sum()
    assert claim_int == cref.mont_to_ints(0, claim.reshape(1, 4))[0]
    prover = zk.SumcheckProver(1)
    proof = prover.prove(pp.clone(), claim_int)
    assert (proof._round_polys_mont == rp).all()
    assert prover.final_evals == cref.mont_to_ints(0, fin)
    assert zk.SumcheckVerifier.verify(pp, proof) is True
    bad = zk.SumcheckProof.from_values(0, claim_int + 1, proof.round_polys)
    with pytest.raises(zk.ZkError, match="claimed_sum"):
        zk.SumcheckVerifier.verify(pp, bad)


def test_prove_host_buffers_e2e(zk, ctx, cref):
    """zk_sumcheck_prove_host: host tables in, proof out (the end-to-end entry bench.py times).  The device landing
    buffers stay with the context between calls: grow, shrink, more and fewer factors, in sequence."""
    import ctypes as C

    from zk_b200 import _ffi

    # 2^21 and 2^22 entries: the sliced upload with round 0 overlapped (api_sumcheck.cu); smaller: upload, then prove
    for n, m, d in [(15, 3, 3), (17, 2, 2), (12, 3, 3), (17, 4, 4), (9, 1, 1), (15, 3, 3), (22, 3, 3), (21, 2, 2), (21, 1, 1), (21, 3, 2)]:
        refs = [cref.gen_table(0, 5 + n, k, n) for k in range(m)]
        claim = cref.product_sum(0, refs, n)
        rp, ch, fin = cref.prove(0, refs, n, d, claim, False, fast=True)
        arr = (C.c_void_p * m)(*[r.ctypes.data for r in refs])
        rp_g = np.zeros_like(rp); ch_g = np.zeros_like(ch); fin_g = np.zeros_like(fin); sum_g = np.zeros(4, dtype=np.uint64)
        st = _ffi.lib().zk_sumcheck_prove_host(ctx.h, 0, arr, m, n, d, None, 0, rp_g.ctypes.data, ch_g.ctypes.data,
                                               fin_g.ctypes.data, sum_g.ctypes.data)
        assert st == 0, (n, m, d)
        assert (sum_g == claim).all() and (rp_g == rp).all() and (ch_g == ch).all() and (fin_g == fin).all(), (n, m, d)
        if n >= 21:  # a caller-supplied claim, wrong on purpose: it only enters the transcript (prover.rs:42)
            bad = cref.ints_to_mont(0, [(cref.mont_to_ints(0, claim.reshape(1, 4))[0] + 5) % O.FIELDS[0].p])[0]
            rp, ch, fin = cref.prove(0, refs, n, d, bad, False, fast=True)
            st = _ffi.lib().zk_sumcheck_prove_host(ctx.h, 0, arr, m, n, d, bad.ctypes.data, 0, rp_g.ctypes.data, ch_g.ctypes.data,
                                                   fin_g.ctypes.data, None)
            assert st == 0 and (rp_g == rp).all() and (ch_g == ch).all() and (fin_g == fin).all(), (n, m, d, "wrong claim")


# ---- properties at sizes the oracle does not reach ------------------------------------------------------------------
@pytest.mark.parametrize("n,m,d", [(24, 2, 2), (24, 3, 3)])
def test_large_prove_self_consistency(zk, ctx, n, m, d):
    """2^24 entries (BASELINE config 2 shape and the degree-3 shape): the verifier's round checks accept,
    the subclaim equals the product of the final folded factors, and equals an independent evaluation of
    the untouched tables at the challenge point (encode -> decode round trip of the whole protocol)."""
    tabs = gpu_tables(zk, 0, O.DEFAULT_SEED, n, m)
    pp = zk.ProductPoly.new(tabs)
    keep = pp.clone()
    claim = pp.sum()
    prover = zk.SumcheckProver(d)
    proof, chs = prover.prove_partial(pp, claim)
    sub = zk.SumcheckVerifier.verify_partial(proof)
    assert sub.challenges == chs
    prod = 1
    for x in prover.final_evals:
        prod = prod * x % F.p
    assert sub.sum == prod
    assert keep.evaluate(chs) == sub.sum
    assert (proof.round_polys[0][0] + proof.round_polys[0][1]) % F.p == claim


def test_config2_bit_exact_vs_cpu_2p20_shape(zk, ctx, cref):
    """Config-2 shape (product of 2 MLEs, D=2) with the full-table absorb, bit-exact vs the CPU transcript."""
    n = 18
    refs = [cref.gen_table(0, O.DEFAULT_SEED, k, n) for k in range(2)]
    claim = cref.product_sum(0, refs, n)
    rp, ch, fin = cref.prove(0, refs, n, 2, claim, True)
    pp = zk.ProductPoly.new(gpu_tables(zk, 0, O.DEFAULT_SEED, n, 2))
    prover = zk.SumcheckProver(2)
    proof = prover.prove(pp.clone(), cref.mont_to_ints(0, claim.reshape(1, 4))[0])
    assert (proof._round_polys_mont == rp).all()
    assert zk.SumcheckVerifier.verify(pp, proof) is True


def test_upload_local_round_trip(zk, ctx, cref):
    """zk_table_upload_local: a shard uploaded as it is comes back unchanged."""
    loc = cref.gen_table(0, 2, 2, 9)
    assert (zk.MultiLinearPolynomial.new_local(9, loc).evaluation_slice_mont() == loc).all()


def test_square_of_one_polynomial_vs_oracle(zk, ctx, cref):
    """ProductPoly::new(vec![f.clone(), f.clone(), g]) (the reference's factors are owned vectors): the mirror clones a handle
    listed twice, the proof equals the oracle's on [f, f, g]; the raw C ABI refuses the same handle twice on the in-place
    paths (the fused fold would fold that buffer once per listing)."""
    from zk_b200 import _ffi

    lib = _ffi.lib()
    n, d, seed = 11, 3, 77
    refs = [cref.gen_table(0, seed, 0, n), cref.gen_table(0, seed, 0, n), cref.gen_table(0, seed, 1, n)]
    claim = cref.product_sum(0, refs, n)
    rp, ch, fin = cref.prove(0, refs, n, d, claim, False)
    f = zk.MultiLinearPolynomial.generate(n, 0, seed=seed)
    g = zk.MultiLinearPolynomial.generate(n, 1, seed=seed)
    pp = zk.ProductPoly.new([f, f, g])
    assert (pp.sum_mont() == claim).all()
    prover = zk.SumcheckProver(d)
    proof, chs = prover.prove_partial(pp, cref.mont_to_ints(0, claim.reshape(1, 4))[0])
    assert (proof._round_polys_mont == rp).all()
    assert chs == cref.mont_to_ints(0, ch) and prover.final_evals == cref.mont_to_ints(0, fin)
    # raw ABI: the same handle twice
    f2 = zk.MultiLinearPolynomial.generate(n, 0, seed=seed)
    arr = zk._table_array([f2, f2])
    out = np.zeros((n, 3, 4), dtype=np.uint64)
    assert lib.zk_sumcheck_prove(ctx.h, arr, 2, 2, claim.ctypes.data, 0, out.ctypes.data, None, None) == 12
    r = zk.to_mont(0, [5])
    assert lib.zk_product_fold_inplace(ctx.h, arr, 2, r.ctypes.data) == 12
    assert lib.zk_product_fold_then_round_poly(ctx.h, arr, 2, 2, r.ctypes.data, out.ctypes.data) == 12
    # read-only calls accept it: sum of f*f
    s = np.zeros(4, dtype=np.uint64)
    assert lib.zk_product_sum(ctx.h, arr, 2, s.ctypes.data) == 0
    assert (s == cref.product_sum(0, refs[:2], n)).all()
    assert (f2.evaluation_slice_mont() == refs[0]).all()  # untouched by the refused calls


def test_randomised_shapes_vs_c_oracle(zk, ctx, cref):
    """A seeded sweep over (field, n, m, D) — 1..8 factors, MAX_VAR_DEGREE 1..5 (5: the generic one-point-per-launch path),
    1..15 variables — so that every kernel the dispatcher can pick for a round (streaming, latency kernel with 2m <= 8
    lanes, the ragged single chunk of tables under 32 items, Toom / plain point sets, derived and summed S(1)) meets the
    oracle on shapes nobody chose by hand.  Claims are true for even cases and false for odd ones."""
    import random

    rng = random.Random(20261018)
    for case in range(40):
        fid = rng.randrange(2)
        m = rng.choice([1, 2, 3, 3, 4, 5, 8])
        d = rng.choice([1, 2, 3, 3, 4, 5])
        n = rng.randrange(1, 16 if m <= 4 else 12)
        seed = rng.randrange(1 << 40)
        refs = [cref.gen_table(fid, seed, k, n) for k in range(m)]
        claim = cref.product_sum(fid, refs, n)
        if case & 1:
            claim = cref.ints_to_mont(fid, [rng.randrange(1 << 200)])[0]
        rp, ch, fin = cref.prove(fid, refs, n, d, claim, False, fast=True)
        pp = zk.ProductPoly.new(gpu_tables(zk, fid, seed, n, m))
        prover = zk.SumcheckProver(d)
        proof, chs = prover.prove_partial(pp, cref.mont_to_ints(fid, claim.reshape(1, 4))[0])
        tag = (case, fid, n, m, d)
        assert (proof._round_polys_mont == rp).all(), tag
        assert chs == cref.mont_to_ints(fid, ch), tag
        assert prover.final_evals == cref.mont_to_ints(fid, fin), tag
