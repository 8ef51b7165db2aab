"""CPU suite: the sum-of-products extension of the oracle, and the host half of its C ABI.

The reference stops at one product (ProductPoly, polynomial/src/product_poly.rs:4-10), so nothing in its tests pins a
sum of products.  What pins it here:
  * a single-term sum is bit-for-bit the reference's ProductPoly proof (same prover loop, prover.rs:33-73);
  * the round polynomial is linear in the terms (round 0 equals the sum of the terms' ProductPoly round polynomials);
  * the reference's own verifier loop (verifier.rs:44-78) accepts the proofs and its sub-claim equals P at the
    challenges, rebuilt from the final table evaluations;
  * two independently written restatements (Python big-int, reference-shaped C) agree, and both reproduce the
    committed golden file tests/golden/sop_vectors.json.
"""
import json
import os

import numpy as np
import pytest

import zkoracle as O
from conftest import ROOT

GKR_TERMS = [[0, 2], [0, 3], [1, 2, 3]]  # add.Wb + add.Wc + mul.Wb.Wc over tables [add, mul, Wb, Wc]
SHAPES = [
    # (field id, n_vars, n_tables, terms, degree)
    (0, 1, 4, GKR_TERMS, 3),
    (0, 2, 4, GKR_TERMS, 3),
    (0, 4, 4, GKR_TERMS, 3),
    (1, 3, 4, GKR_TERMS, 3),
    (0, 3, 3, [[0], [1, 2]], 2),            # a linear term next to a quadratic one
    (0, 3, 2, [[0, 0], [1]], 2),            # a squared table
    (1, 4, 5, [[0, 1, 2, 3], [4], [2, 4]], 4),
    (0, 3, 4, GKR_TERMS, 2),                # MAX_VAR_DEGREE below the true degree: not validated, like the reference
]


def tables_int(field_id, n_tables, n, seed=O.DEFAULT_SEED):
    F = O.FIELDS[field_id]
    return [O.gen_table(F, seed, 20 + k, n) for k in range(n_tables)]


def py_sop(field_id, n_tables, n, terms):
    F = O.FIELDS[field_id]
    return O.SumOfProductsPoly([O.MultiLinearPolynomial(F, n, t) for t in tables_int(field_id, n_tables, n)], terms)


@pytest.fixture(scope="module")
def sop_golden():
    with open(os.path.join(ROOT, "tests", "golden", "sop_vectors.json")) as f:
        return json.load(f)


def test_single_term_is_the_product_poly_proof():
    F = O.BLS12_381_FR
    for n, m, d in [(3, 3, 3), (4, 2, 2), (2, 1, 1)]:
        tabs = [O.MultiLinearPolynomial(F, n, O.gen_table(F, 7, k, n)) for k in range(m)]
        pp = O.ProductPoly(tabs)
        sp = O.SumOfProductsPoly(tabs, [list(range(m))])
        claim = sum(pp.prod_reduce()) % F.p
        assert sum(sp.prod_reduce()) % F.p == claim
        a, ca = O.SumcheckProver(d).prove_partial(pp, claim)
        b, cb = O.SumcheckProver(d).prove_partial(sp, claim)
        assert a.round_polys == b.round_polys and ca == cb
        assert O.SumcheckProver(d).prove(pp.clone(), claim).round_polys == O.SumcheckProver(d).prove(sp.clone(), claim).round_polys


def test_round_polynomial_is_linear_in_the_terms():
    F = O.BLS12_381_FR
    n = 4
    sp = py_sop(0, 4, n, GKR_TERMS)
    want = [0] * 4
    for term in GKR_TERMS:
        pp = O.ProductPoly([sp.polynomials[k] for k in term])
        for t in range(4):
            want[t] = (want[t] + sum(pp.partial_evaluate(0, [t]).prod_reduce())) % F.p
    got = [sum(sp.partial_evaluate(0, [t]).prod_reduce()) % F.p for t in range(4)]
    assert got == want


def test_distributive_form_of_the_gkr_layer():
    """add.(Wb + Wc) + mul.Wb.Wc, evaluated the way a GKR prover writes it, is the three-term sum."""
    F = O.BLS12_381_FR
    n = 3
    sp = py_sop(0, 4, n, GKR_TERMS)
    add, mul, wb, wc = [q.evaluations for q in sp.polynomials]
    direct = [(a * (b + c) + m * b * c) % F.p for a, m, b, c in zip(add, mul, wb, wc)]
    assert sp.prod_reduce() == direct
    pt = [5, F.p - 3, 1 << 200]
    vals = [q.evaluate(pt) for q in sp.polynomials]
    assert sp.evaluate(pt) == (vals[0] * (vals[2] + vals[3]) + vals[1] * vals[2] * vals[3]) % F.p


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"f{s[0]}_n{s[1]}_k{s[2]}_t{len(s[3])}_d{s[4]}")
def test_python_and_c_oracles_agree_and_verifier_accepts(cref, shape):
    fid, n, nt, terms, d = shape
    F = O.FIELDS[fid]
    sp = py_sop(fid, nt, n, terms)
    claim = sum(sp.prod_reduce()) % F.p
    prover = O.SumcheckProver(d)
    proof, ch = prover.prove_partial(sp.clone(), claim)
    finals = [q.evaluations[0] for q in prover.final_poly.polynomials]
    # C restatement
    tabs = [cref.ints_to_mont(fid, t) for t in tables_int(fid, nt, n)]
    csum = cref.sop_sum(fid, tabs, terms, n)
    assert cref.mont_to_ints(fid, csum.reshape(1, 4))[0] == claim
    rp, cch, cfin = cref.prove_sop(fid, tabs, terms, n, d, csum)
    assert cref.mont_to_ints(fid, rp.reshape(-1, 4)) == [x for r in proof.round_polys for x in r]
    assert cref.mont_to_ints(fid, cch) == ch
    assert cref.mont_to_ints(fid, cfin) == finals
    # the reference's verifier loop; only proofs whose MAX_VAR_DEGREE covers the true degree pass the round checks
    true_degree = max(len(t) for t in terms)
    if d >= true_degree:
        sub = O.SumcheckVerifier.verify_partial(F, proof)
        assert sub.challenges == ch
        assert sub.sum == sp.combine(finals) == sp.evaluate(ch)
        bad = O.SumcheckProof((claim + 1) % F.p, proof.round_polys)
        with pytest.raises(O.OracleError, match="claimed_sum != p\\(0\\) \\+ p\\(1\\)"):
            O.SumcheckVerifier.verify_partial(F, bad)
    # with the initial absorb (tables in order)
    rp2, ch2, _ = cref.prove_sop(fid, tabs, terms, n, d, csum, absorb=True)
    p2 = O.SumcheckProver(d).prove(sp.clone(), claim)
    assert cref.mont_to_ints(fid, rp2.reshape(-1, 4)) == [x for r in p2.round_polys for x in r]
    assert (ch2 != cch).any()


def test_python_and_c_oracles_agree_on_random_term_structures(cref):
    rng = np.random.default_rng(7)
    for trial in range(25):
        fid = int(rng.integers(0, 2))
        F = O.FIELDS[fid]
        nt = int(rng.integers(1, 7))
        terms = [[int(x) for x in rng.integers(0, nt, size=int(rng.integers(1, 6)))] for _ in range(int(rng.integers(1, 7)))]
        n, d = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        ints = [O.gen_table(F, 500 + trial, 20 + k, n) for k in range(nt)]
        sp = O.SumOfProductsPoly([O.MultiLinearPolynomial(F, n, t) for t in ints], terms)
        claim = sum(sp.prod_reduce()) % F.p
        proof, ch = O.SumcheckProver(d).prove_partial(sp, claim)
        tabs = [cref.ints_to_mont(fid, t) for t in ints]
        csum = cref.sop_sum(fid, tabs, terms, n)
        assert cref.mont_to_ints(fid, csum.reshape(1, 4))[0] == claim, (trial, terms)
        rp, cch, _ = cref.prove_sop(fid, tabs, terms, n, d, csum)
        assert cref.mont_to_ints(fid, rp.reshape(-1, 4)) == [x for r in proof.round_polys for x in r], (trial, terms, d)
        assert cref.mont_to_ints(fid, cch) == ch


def test_c_oracle_single_term_equals_reference_shaped_product_prover(cref):
    for fid, n, m, d in [(0, 5, 3, 3), (1, 4, 2, 2)]:
        tabs = [cref.gen_table(fid, 11, k, n) for k in range(m)]
        claim = cref.product_sum(fid, tabs, n)
        a = cref.prove(fid, tabs, n, d, claim, False)
        b = cref.prove_sop(fid, tabs, [list(range(m))], n, d, claim)
        for x, y in zip(a, b):
            assert (x == y).all()


def test_oracles_reproduce_the_sop_golden_file(cref, sop_golden):
    assert len(sop_golden["cases"]) >= 6
    for case in sop_golden["cases"]:
        fid, n, nt, terms, d = case["field"], case["n_vars"], case["n_tables"], case["terms"], case["degree"]
        tabs = [cref.gen_table(fid, case["seed"], 20 + k, n) for k in range(nt)]
        csum = cref.sop_sum(fid, tabs, terms, n)
        assert "%064x" % cref.mont_to_ints(fid, csum.reshape(1, 4))[0] == case["sum"]
        rp, ch, fin = cref.prove_sop(fid, tabs, terms, n, d, csum)
        assert ["%064x" % v for v in cref.mont_to_ints(fid, rp.reshape(-1, 4))] == case["round_polys"]
        assert ["%064x" % v for v in cref.mont_to_ints(fid, ch)] == case["challenges"]
        assert ["%064x" % v for v in cref.mont_to_ints(fid, fin)] == case["final_evals"]


# ---- host half of the C ABI (no GPU needed) -------------------------------------------------------------
def test_abi_sop_combine_vs_oracle(zk):
    import ctypes as C

    for fid in (0, 1):
        F = O.FIELDS[fid]
        vals = [O.gen_element(3, 9, i) % F.p for i in range(4)]
        mont = zk.to_mont(fid, vals)
        tl = np.array([len(t) for t in GKR_TERMS], dtype=np.uint8)
        tf = np.array([k for t in GKR_TERMS for k in t], dtype=np.uint8)
        out = np.zeros(4, dtype=np.uint64)
        st = zk.lib().zk_sop_combine(fid, tl.ctypes.data, tf.ctypes.data, len(GKR_TERMS), mont.ctypes.data, 4, out.ctypes.data)
        assert st == 0
        want = (vals[0] * (vals[2] + vals[3]) + vals[1] * vals[2] * vals[3]) % F.p
        assert zk.from_mont(fid, out)[0] == want
        # a factor index past n_tables is refused
        assert zk.lib().zk_sop_combine(fid, tl.ctypes.data, tf.ctypes.data, len(GKR_TERMS), mont.ctypes.data, 3, out.ctypes.data) == 12


def test_mirror_sop_argument_checks_need_no_gpu(zk):
    with pytest.raises(zk.ZkError, match="cannot create product polynomial from empty polynomials"):
        zk.SumOfProductsPoly.new([], [[0]])


def test_cpp_mirror_of_the_sop_entry_points_compiles_and_runs_its_host_part():
    import subprocess

    so_dir = os.path.join(ROOT, "zk_b200")
    exe = os.path.join(ROOT, "build", "test_sop_mirror_compiles")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_sop_mirror_compiles.cpp"), "-L", so_dir, "-lzk_b200", f"-Wl,-rpath,{so_dir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "SOP MIRROR OK" in r.stdout, r.stdout + r.stderr


def test_sop_kernel_source_replayed_on_the_host(tmp_path):
    """zk_b200/csrc/sop_kernel.cuh compiled as plain C++ and run thread by thread against a naive model
    (tests/cpp/test_sop_kernel_host.cpp): index conventions of the fused in-place fold, term bookkeeping, ragged sizes."""
    import subprocess

    exe = str(tmp_path / "test_sop_kernel_host")
    cmd = ["g++", "-std=c++17", "-O2", "-I", "/usr/local/cuda/include", "-I", os.path.join(ROOT, "zk_b200", "csrc"),
           os.path.join(ROOT, "tests", "cpp", "test_sop_kernel_host.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300).stdout
    assert "2884 cases, 0 mismatches, 0 guard hits" in out, out
