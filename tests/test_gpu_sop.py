"""GPU suite: the sum-of-products sumcheck (P = sum_t prod_{k in term t} A_k, the GKR layer shape
add.Wb + add.Wc + mul.Wb.Wc) through the C ABI, bit-exact against the oracle, the committed golden file
(tests/golden/sop_vectors.json), and the already-validated ProductPoly kernels (a single-term sum IS the product
proof; round 0 is linear in the terms)."""
import json
import os

import numpy as np
import pytest

import zkoracle as O
from conftest import ROOT

pytestmark = pytest.mark.gpu

GKR_TERMS = [[0, 2], [0, 3], [1, 2, 3]]


def gpu_sop(zk, fid, seed, n, nt, terms):
    return zk.SumOfProductsPoly.new([zk.MultiLinearPolynomial.generate(n, 20 + k, seed=seed, field=fid) for k in range(nt)], terms)


def first_mismatch(a, b):
    a, b = np.asarray(a).reshape(-1, 4), np.asarray(b).reshape(-1, 4)
    bad = np.nonzero((a != b).any(axis=1))[0]
    return None if bad.size == 0 else (int(bad[0]), int(bad.size), a.shape[0])


def test_sop_golden_file(zk, ctx):
    with open(os.path.join(ROOT, "tests", "golden", "sop_vectors.json")) as f:
        cases = json.load(f)["cases"]
    for case in cases:
        fid, n, nt, terms, d = case["field"], case["n_vars"], case["n_tables"], case["terms"], case["degree"]
        tag = (fid, n, nt, terms, d)
        sp = gpu_sop(zk, fid, case["seed"], n, nt, terms)
        claim = sp.sum()
        assert "%064x" % claim == case["sum"], tag
        assert ["%064x" % v for v in sp.round_poly(d)] == case["round_polys"][: d + 1], tag
        keep = sp.clone()
        prover = zk.SumcheckProver(d)
        proof, ch = prover.prove_partial(sp, claim)
        got = ["%064x" % x for r in proof.round_polys for x in r]
        bad = [i for i, (a, b) in enumerate(zip(got, case["round_polys"])) if a != b]
        assert not bad, (tag, "first differing evaluation (round, t):", divmod(bad[0], d + 1), len(bad))
        assert ["%064x" % c for c in ch] == case["challenges"], tag
        assert ["%064x" % v for v in prover.final_evals] == case["final_evals"], tag
        if d >= max(len(t) for t in terms):
            sub = zk.SumcheckVerifier.verify_partial(proof)
            assert sub.challenges == ch
            assert sub.sum == keep.combine(prover.final_evals) == keep.evaluate(ch), tag


@pytest.mark.parametrize("fid,n,d", [(0, 12, 3), (1, 13, 3), (0, 14, 2), (0, 19, 3)])
def test_gkr_shape_vs_c_oracle(zk, ctx, cref, fid, n, d):
    """n = 19: 2^18 items in round 0, more than one grid of blocks -> the grid-stride loop and the multi-block
    reduction; the fused fold+sum launches run for table sizes 2^19 .. 4."""
    seed = 4242 + n
    sp = gpu_sop(zk, fid, seed, n, 4, GKR_TERMS)
    refs = [cref.gen_table(fid, seed, 20 + k, n) for k in range(4)]
    rsum = cref.sop_sum(fid, refs, GKR_TERMS, n)
    assert (sp.sum_mont() == rsum).all()
    rp, ch, fin = cref.prove_sop(fid, refs, GKR_TERMS, n, d, rsum)
    prover = zk.SumcheckProver(d)
    proof, gch = prover.prove_partial(sp, zk.from_mont(fid, rsum)[0])
    assert first_mismatch(proof._round_polys_mont, rp) is None, ("round polys (index, count, total)", first_mismatch(proof._round_polys_mont, rp))
    assert gch == cref.mont_to_ints(fid, ch)
    assert prover.final_evals == cref.mont_to_ints(fid, fin)


@pytest.mark.parametrize("m,d,n", [(3, 3, 11), (2, 2, 12), (1, 1, 10), (4, 4, 9)])
def test_single_term_equals_product_poly_on_the_gpu(zk, ctx, m, d, n):
    tabs = [zk.MultiLinearPolynomial.generate(n, k, seed=31) for k in range(m)]
    pp = zk.ProductPoly.new([t.clone() for t in tabs])
    sp = zk.SumOfProductsPoly.new(tabs, [list(range(m))])
    claim = pp.sum()
    assert sp.sum() == claim
    assert sp.round_poly(d) == pp.round_poly(d)
    a, ca = zk.SumcheckProver(d).prove_partial(pp, claim)
    pb = zk.SumcheckProver(d)
    b, cb = pb.prove_partial(sp, claim)
    assert a.round_polys == b.round_polys and ca == cb


def test_round_zero_is_linear_in_the_terms(zk, ctx):
    n, p = 12, O.BLS12_381_FR.p
    sp = gpu_sop(zk, 0, 77, n, 4, GKR_TERMS)
    want = [0] * 4
    for term in GKR_TERMS:
        rp = zk.ProductPoly.new([sp.polynomials[k] for k in term]).round_poly(3)
        want = [(w + x) % p for w, x in zip(want, rp)]
    assert sp.round_poly(3) == want
    assert sp.sum() == (want[0] + want[1]) % p


def test_repeated_factor_and_many_tables(zk, ctx, cref):
    for fid, n, nt, terms, d in [(0, 9, 2, [[0, 0], [1]], 2), (1, 8, 5, [[0, 1, 2, 3], [4], [2, 4]], 4),
                                 (0, 10, 8, [[0, 1], [2, 3], [4, 5], [6, 7], [0, 7], [1, 6], [2, 5], [3, 4]], 2)]:
        sp = gpu_sop(zk, fid, 5, n, nt, terms)
        refs = [cref.gen_table(fid, 5, 20 + k, n) for k in range(nt)]
        rsum = cref.sop_sum(fid, refs, terms, n)
        rp, ch, fin = cref.prove_sop(fid, refs, terms, n, d, rsum)
        prover = zk.SumcheckProver(d)
        proof, gch = prover.prove_partial(sp, zk.from_mont(fid, rsum)[0])
        assert (proof._round_polys_mont == rp).all(), (fid, n, nt, terms, d)
        assert prover.final_evals == cref.mont_to_ints(fid, fin)


def test_prove_with_initial_absorb_and_wrong_claim(zk, ctx, cref):
    fid, n, d = 0, 7, 3
    refs = [cref.gen_table(fid, 3, 20 + k, n) for k in range(4)]
    rsum = cref.sop_sum(fid, refs, GKR_TERMS, n)
    rp, ch, fin = cref.prove_sop(fid, refs, GKR_TERMS, n, d, rsum, absorb=True)
    proof = zk.SumcheckProver(d).prove(gpu_sop(zk, fid, 3, n, 4, GKR_TERMS), zk.from_mont(fid, rsum)[0])
    assert (proof._round_polys_mont == rp).all()
    # a wrong claim only enters the transcript: S(0), S(1) stay the true values (sumcheck/src/lib.rs:115-122)
    wrong = cref.ints_to_mont(fid, [12345])[0]
    rp2, _, _ = cref.prove_sop(fid, refs, GKR_TERMS, n, d, wrong)
    proof2, _ = zk.SumcheckProver(d).prove_partial(gpu_sop(zk, fid, 3, n, 4, GKR_TERMS), 12345)
    assert (proof2._round_polys_mont == rp2).all()
    with pytest.raises(zk.ZkError, match="claimed_sum != p\\(0\\) \\+ p\\(1\\)"):
        zk.SumcheckVerifier.verify_partial(proof2)


def test_zero_variables_and_argument_errors(zk, ctx):
    vals = [3, 5, 7, 11]
    tabs = [zk.MultiLinearPolynomial.new(0, [v]) for v in vals]
    sp = zk.SumOfProductsPoly.new(tabs, GKR_TERMS)
    assert sp.sum() == 3 * (7 + 11) + 5 * 7 * 11
    prover = zk.SumcheckProver(3)
    proof, ch = prover.prove_partial(sp, sp.sum())
    assert proof.round_polys == [] and ch == [] and prover.final_evals == vals
    t = [zk.MultiLinearPolynomial.generate(4, k) for k in range(2)]
    with pytest.raises(zk.ZkError) as e:  # the same table twice in tables[]
        zk.SumOfProductsPoly([t[0], t[0]], [[0, 1]]).round_poly(2)
    assert e.value.status == 12
    with pytest.raises(zk.ZkError) as e:  # MAX_VAR_DEGREE outside 1..4
        zk.SumOfProductsPoly.new(t, [[0, 1]]).round_poly(5)
    assert e.value.status == 13
    with pytest.raises(zk.ZkError) as e:
        zk.SumOfProductsPoly(t, [[0, 2]]).round_poly(2)
    assert e.value.status == 12
    with pytest.raises(zk.ZkError, match="evaluate must assign to all variables"):
        zk.SumOfProductsPoly.new(t, [[0, 1]]).evaluate([1, 2, 3])
    u = zk.MultiLinearPolynomial.generate(5, 0)
    with pytest.raises(zk.ZkError, match="don't share the same number of variables"):
        zk.SumOfProductsPoly.new([t[0], u], [[0, 1]])


def test_prove_and_verify_with_the_initial_absorb(zk, ctx):
    """prove (tables absorbed first) <-> SumcheckVerifier.verify over the sum of products (zk_sumcheck_verify_sop):
    accept, Ok(false) is unreachable without breaking a round check first, the reference's Err strings."""
    sp = gpu_sop(zk, 0, 11, 9, 4, GKR_TERMS)
    keep = sp.clone()
    proof = zk.SumcheckProver(3).prove(sp, keep.sum())
    assert zk.SumcheckVerifier.verify(keep, proof) is True
    with pytest.raises(zk.ZkError, match="require 1 round poly for each variable"):
        zk.SumcheckVerifier.verify(keep, zk.SumcheckProof.from_values(0, proof.sum, proof.round_polys[:-1]))
    with pytest.raises(zk.ZkError, match="claimed_sum != p\\(0\\) \\+ p\\(1\\)"):
        zk.SumcheckVerifier.verify(keep, zk.SumcheckProof.from_values(0, proof.sum + 1, proof.round_polys))


def test_randomised_term_structures_vs_c_oracle(zk, ctx, cref):
    """A seeded sweep over term structures: 1..6 tables, 1..5 terms of 1..4 factors (repeats allowed, every table used),
    MAX_VAR_DEGREE 1..4 at, above and below the longest term, 2..13 variables — so that the launcher's common-factor
    extraction (terms that differ in one table), the Toom point set (degree 3, no term longer than three), the latency
    kernel (at most four tables) and the streaming kernel all meet the oracle on shapes nobody chose by hand."""
    import random

    rng = random.Random(4242)
    for case in range(30):
        fid = rng.randrange(2)
        nt = rng.randrange(1, 7)
        n_terms = rng.randrange(1, 6)
        terms = [[rng.randrange(nt) for _ in range(rng.randrange(1, 5))] for _ in range(n_terms)]
        if case % 3 == 0 and nt >= 3:  # plant a pair that differs in exactly one table: x.a + x.b
            terms[0] = [0, 1]
            terms.append([0, 2])
        for k in range(nt):  # every table appears somewhere (its fold is part of the proof's final evaluations either way)
            if not any(k in t for t in terms):
                terms[rng.randrange(len(terms))][0:0] = [k] if len(terms[0]) < 4 else []
        terms = [t[:4] for t in terms][:8]
        d = rng.choice([1, 2, 3, 3, 4])
        n = rng.randrange(2, 14)
        seed = rng.randrange(1 << 40)
        sp = gpu_sop(zk, fid, seed, n, nt, terms)
        refs = [cref.gen_table(fid, seed, 20 + k, n) for k in range(nt)]
        rsum = cref.sop_sum(fid, refs, terms, n)
        assert (sp.sum_mont() == rsum).all(), (case, terms)
        rp, ch, fin = cref.prove_sop(fid, refs, terms, n, d, rsum)
        prover = zk.SumcheckProver(d)
        proof, gch = prover.prove_partial(sp, zk.from_mont(fid, rsum)[0])
        tag = (case, fid, n, nt, terms, d)
        assert first_mismatch(proof._round_polys_mont, rp) is None, (tag, first_mismatch(proof._round_polys_mont, rp))
        assert gch == cref.mont_to_ints(fid, ch), tag
        assert prover.final_evals == cref.mont_to_ints(fid, fin), tag
