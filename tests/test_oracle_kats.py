"""CPU suite, part 1: pins the oracle (Python big-int restatement and C restatement) against every
known-answer test the reference holds for the path, and against the committed golden vectors.
Reference test names are kept (file:line in each docstring)."""
import numpy as np
import pytest

import zkoracle as O
from conftest import hx

F = O.BLS12_381_FR
Fr = lambda v: v % F.p


def mle(vals, n=None, field=F):
    n = (len(vals).bit_length() - 1) if n is None else n
    return O.MultiLinearPolynomial(field, n, [v % field.p for v in vals])


# ---- polynomial/src/multilinear/pairing_index.rs:32-97 ------------------------------------------------
def test_insert_bit():
    assert O.insert_bit(0b1011, 1, 0) == 0b10101
    assert O.insert_bit(0b1011, 4, 1) == 0b11011
    assert O.insert_bit(0, 0, 1) == 1


def test_index_pair():
    assert list(O.index_pair(1, 0)) == [(0, 1)]
    assert list(O.index_pair(2, 0)) == [(0, 2), (1, 3)]
    assert list(O.index_pair(2, 1)) == [(0, 1), (2, 3)]
    assert list(O.index_pair(3, 0)) == [(0, 4), (1, 5), (2, 6), (3, 7)]
    assert list(O.index_pair(3, 1)) == [(0, 2), (1, 3), (4, 6), (5, 7)]
    assert list(O.index_pair(3, 2)) == [(0, 1), (2, 3), (4, 5), (6, 7)]


# ---- polynomial/src/multilinear/evaluation_form.rs:112-202 ---------------------------------------------
def test_new_multilinear_poly():
    with pytest.raises(O.OracleError, match="evaluation vec len should equal 2\\^n_vars"):
        mle([3, 1, 2], 2)
    with pytest.raises(O.OracleError):
        mle([3, 1], 2)
    mle([3, 1], 1)
    mle([3, 1, 2, 5], 2)


def test_partial_evaluate_single_variable():
    poly = mle([3, 1, 2, 5])
    assert poly.partial_evaluate(0, [5]).evaluations == [Fr(-2), Fr(21)]
    assert poly.partial_evaluate(0, [0]).evaluations == [3, 1]


def test_partial_evaluate_consecutive_variables():
    poly = mle([0, 0, 0, 3, 0, 0, 2, 5])
    out = poly.partial_evaluate(1, [2, 3]).evaluations
    assert out == [18, 22]


def test_full_evaluation():
    poly = mle([0, 0, 0, 3, 0, 0, 2, 5])
    assert poly.evaluate([2, 3, 4]) == 48
    with pytest.raises(O.OracleError, match="evaluate must assign to all variables"):
        poly.evaluate([2, 3])


# ---- polynomial/src/product_poly.rs:97-196 ---------------------------------------------------------------
def test_product_poly_creation():
    with pytest.raises(O.OracleError, match="empty polynomials"):
        O.ProductPoly([])
    with pytest.raises(O.OracleError, match="share the same number of variables"):
        O.ProductPoly([mle([1, 2]), mle([1, 2, 3, 4])])
    O.ProductPoly([mle([1, 2, 3, 4]), mle([1, 2, 3, 4])])


def test_product_poly_evaluate_and_partial():
    p1, p2 = mle([0, 0, 0, 3, 0, 0, 2, 5]), mle([1, 2, 3, 4, 5, 6, 7, 8])
    pp = O.ProductPoly([p1, p2])
    pt = [2, 3, 4]
    assert pp.evaluate(pt) == (p1.evaluate(pt) * p2.evaluate(pt)) % F.p
    pe = pp.partial_evaluate(0, [7])
    assert pe.polynomials[0] == p1.partial_evaluate(0, [7]) and pe.polynomials[1] == p2.partial_evaluate(0, [7])


def test_prod_reduce():
    pp = O.ProductPoly([mle([2, 8, 10, 14]), mle([2, 8, 10, 22])])
    assert pp.prod_reduce() == [4, 64, 100, 308]


# ---- sumcheck/src/lib.rs:53-122 ----------------------------------------------------------------------------
def p_2ab_3bc():
    return mle([0, 0, 0, 3, 0, 0, 2, 5])  # coefficient_form.rs:1321-1347 fixture


def test_sumcheck_correct_sum_multilinear():
    pp = O.ProductPoly([p_2ab_3bc()])
    proof = O.SumcheckProver(1).prove(pp.clone(), 10)
    assert O.SumcheckVerifier.verify(pp, proof) is True


def test_correct_sum_multivariate_deg_2():
    pp = O.ProductPoly([mle([3, 3, 5, 5]), mle([0, 0, 0, 1])])
    proof = O.SumcheckProver(2).prove(pp.clone(), 5)
    assert O.SumcheckVerifier.verify(pp, proof) is True


def test_correct_sum_prove_partial():
    pp = O.ProductPoly([p_2ab_3bc()])
    proof, _ = O.SumcheckProver(1).prove_partial(pp.clone(), 10)
    sub = O.SumcheckVerifier.verify_partial(F, proof)
    assert pp.evaluate(sub.challenges) == sub.sum


def test_invalid_sum():
    pp = O.ProductPoly([p_2ab_3bc()])
    proof = O.SumcheckProver(1).prove(pp.clone(), 12)
    with pytest.raises(O.OracleError, match="claimed_sum != p\\(0\\) \\+ p\\(1\\)"):
        O.SumcheckVerifier.verify(pp, proof)


def test_verify_wrong_round_count_and_false():
    pp = O.ProductPoly([p_2ab_3bc()])
    proof = O.SumcheckProver(1).prove(pp.clone(), 10)
    short = O.SumcheckProof(proof.sum, proof.round_polys[:2])
    with pytest.raises(O.OracleError, match="require 1 round poly"):
        O.SumcheckVerifier.verify(pp, short)
    other = O.ProductPoly([mle([0, 0, 0, 3, 0, 0, 2, 6])])  # proof of another poly: transcript differs
    with pytest.raises(O.OracleError):
        O.SumcheckVerifier.verify(other, proof)


# ---- univariate (verifier side; polynomial/src/univariate_poly.rs tests are on F17) -----------------------------
def test_univariate_f17():
    f = O.F17
    p = O.UnivariatePolynomial(f, [5, 2, 3])
    assert p.evaluate(2) == (5 + 4 + 12) % 17
    q = O.UnivariatePolynomial.interpolate_xy(f, [0, 1, 2], [p.evaluate(0), p.evaluate(1), p.evaluate(2)])
    assert q.coefficients == [5, 2, 3]
    assert O.UnivariatePolynomial.interpolate(f, [0, 2]).coefficients == [0, 2]


# ---- fft/src/lib.rs:78-82 ------------------------------------------------------------------------------------------
def test_fft():
    a = [0, 2, 34, 3434]
    assert O.ifft(O.BLS12_377_FR, O.fft(O.BLS12_377_FR, a)) == a
    with pytest.raises(ValueError):
        O.fft_internal(O.BLS12_377_FR, [1, 2, 3], 5)


# ---- golden vectors: Python oracle reproduces the committed file; C oracle agrees ---------------------------------------
def test_keccak_kats(golden, cref):
    g = golden["transcript"]
    for impl in (O.keccak256, cref.keccak256):
        assert impl(b"").hex() == g["keccak_empty"]
        assert impl(b"abc").hex() == g["keccak_abc"]
        assert impl(b"a" * 200).hex() == g["keccak_200x61"]
        assert impl(bytes(136)).hex() == g["keccak_136x00"]


def test_generator_anchors(golden, cref):
    g = golden["generator"]
    assert O.gen_element(O.DEFAULT_SEED, 0, 0) == hx(g["elem_k0_i0"])
    assert O.gen_element(O.DEFAULT_SEED, 2, 2**30 - 1) == hx(g["elem_k2_i2p30m1"])
    t = cref.gen_table(0, O.DEFAULT_SEED, 0, 4)
    assert cref.mont_to_ints(0, t)[:2] == [hx(g["elem_k0_i0"]), hx(g["elem_k0_i1"])]
    assert cref.limbs_to_ints(t[:1])[0] == hx(g["elem_k0_i0_mont"])
    t2 = cref.gen_table(0, O.DEFAULT_SEED, 2, 30, first=2**30 - 1, count=1)
    assert cref.mont_to_ints(0, t2)[0] == hx(g["elem_k2_i2p30m1"])
    t3 = cref.gen_table(1, O.DEFAULT_SEED, 1, 4, first=5, count=1)
    assert cref.mont_to_ints(1, t3)[0] == hx(g["elem377_k1_i5"])
    # strided generation == slicing the full table
    full = cref.gen_table(0, 99, 1, 6)
    assert (cref.gen_table(0, 99, 1, 6, first=3, stride=4, count=16) == full[3::4]).all()


@pytest.mark.parametrize("group", ["small_cases", "seeded_cases"])
def test_c_oracle_reproduces_golden_proofs(golden, cref, group):
    for c in golden[group]:
        fid = c["field"]
        if "tables" in c:
            tabs = [cref.ints_to_mont(fid, t) for t in c["tables"]]
            n = len(c["tables"][0]).bit_length() - 1
        else:
            n = c["n"]
            tabs = [cref.gen_table(fid, hx(c["seed"]), k, n) for k in range(c["m"])]
        claim = cref.ints_to_mont(fid, [hx(c["claim"])])[0]
        for fast in (False, True):  # the streamlined prover too, with and without the initial-poly absorb
            rp, ch, fin = cref.prove(fid, tabs, n, c["degree"], claim, c["absorb"], fast=fast)
            assert cref.mont_to_ints(fid, rp.reshape(-1, 4)) == [hx(x) for r in c["round_polys"] for x in r], c["name"]
            assert cref.mont_to_ints(fid, ch) == [hx(x) for x in c["challenges"]], c["name"]
            assert cref.mont_to_ints(fid, fin) == [hx(x) for x in c["finals"]], c["name"]
        if "seed" in c:  # honest claim: C product_sum agrees
            assert cref.mont_to_ints(fid, cref.product_sum(fid, tabs, n).reshape(1, 4))[0] == hx(c["claim"])


def test_c_oracle_verifier(golden, cref):
    by = {c["name"]: c for c in golden["small_cases"] + golden["seeded_cases"]}
    for name in ("B1_prove_2ab3bc_sum10", "B3_prove_deg2_sum5", "seeded_n10_m3_d3_prove", "seeded_n7_m2_d3_prove"):
        c = by[name]
        fid = c["field"]
        if "tables" in c:
            tabs = [cref.ints_to_mont(fid, t) for t in c["tables"]]
        else:
            tabs = [cref.gen_table(fid, hx(c["seed"]), k, c["n"]) for k in range(c["m"])]
        init = b"".join(cref.to_bytes(fid, t) for t in tabs)
        rp = cref.ints_to_mont(fid, [hx(x) for r in c["round_polys"] for x in r]).reshape(len(c["round_polys"]), -1, 4)
        rc, sub, ch = cref.verify_internal(fid, cref.ints_to_mont(fid, [hx(c["claim"])])[0], rp, init)
        assert rc == 0
        assert cref.mont_to_ints(fid, ch) == [hx(x) for x in c["challenges"]]
        prod = 1
        for x in c["finals"]:
            prod = prod * hx(x) % O.FIELDS[fid].p
        assert cref.mont_to_ints(fid, sub.reshape(1, 4))[0] == prod
    c = by["B4_prove_2ab3bc_wrong_sum12"]
    rp = cref.ints_to_mont(0, [hx(x) for r in c["round_polys"] for x in r]).reshape(3, 2, 4)
    init = cref.to_bytes(0, cref.ints_to_mont(0, c["tables"][0]))
    rc, _, _ = cref.verify_internal(0, cref.ints_to_mont(0, [12])[0], rp, init)
    assert rc == 3


def test_c_oracle_partial_evaluate(golden, cref):
    for c in golden["partial_evaluate"]:
        tab = cref.gen_table(0, c["seed"], 0, c["n"])
        assert cref.mont_to_ints(0, tab) == O.gen_table(F, c["seed"], 0, c["n"])
        out = cref.partial_evaluate(0, tab, c["n"], c["initial_var"], cref.ints_to_mont(0, c["assignments"]))
        assert cref.mont_to_ints(0, out) == [hx(x) for x in c["out"]]


def test_c_oracle_fft(golden, cref):
    for c in golden["fft"]:
        a = cref.ints_to_mont(c["field"], [hx(x) for x in c["in"]])
        for fast in (False, True):
            fw = cref.fft(c["field"], a, c["log_n"], fast=fast)
            assert cref.mont_to_ints(c["field"], fw) == [hx(x) for x in c["fft"]]
            assert (cref.fft(c["field"], fw, c["log_n"], inverse=True, fast=fast) == a).all()
    with pytest.raises(ValueError):
        cref.fft(0, np.zeros((1, 4), dtype=np.uint64), 33)


def test_python_oracle_reproduces_golden(golden):
    c = {x["name"]: x for x in golden["seeded_cases"]}["seeded_n7_m2_d3_partial"]
    tabs = [O.gen_table(F, hx(c["seed"]), k, c["n"]) for k in range(c["m"])]
    pp = O.ProductPoly([O.MultiLinearPolynomial(F, c["n"], t) for t in tabs])
    proof, ch = O.SumcheckProver(c["degree"]).prove_partial(pp, hx(c["claim"]))
    assert [hex(x) for x in ch] == c["challenges"]
    assert [[hex(x) for x in r] for r in proof.round_polys] == c["round_polys"]
    g = golden["transcript"]
    tr = O.Transcript(); tr.append(b"zk-b200 transcript golden"); tr.append(bytes(range(200)))
    got = [hex(tr.sample_field_element(F)) for _ in range(3)]
    tr.append(b"more"); got.append(hex(tr.sample_field_element(O.BLS12_377_FR)))
    assert got == g["challenges"]


def test_field_constants():
    for FF, inv64 in ((O.BLS12_381_FR, 0xFFFFFFFEFFFFFFFF), (O.BLS12_377_FR, 0x0A117FFFFFFFFFFF)):
        assert (-pow(FF.p, -1, 1 << 64)) % (1 << 64) == inv64
        assert (-pow(FF.p, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF  # m = -t0 shortcut used by the device multiplier
        assert pow(FF.generator, (FF.p - 1) // 2, FF.p) == FF.p - 1
    assert O.BLS12_381_FR.two_adicity == 32 and O.BLS12_377_FR.two_adicity == 47


def test_fullsize_digest_file_is_the_oracles(cref):
    """tests/golden/fullsize_digests.json (checked by the GPU suite at full size) is reproducible from the C oracle:
    recompute its 2^20 case here (0.3 s); the larger cases come from the same script (make_fullsize_digests.py)."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cases = json.load(open(os.path.join(root, "tests", "golden", "fullsize_digests.json")))["cases"]
    assert {(c["log_n"], c["m"], c["degree"]) for c in cases} >= {(20, 1, 1), (24, 2, 2), (26, 3, 3)}
    c = next(c for c in cases if c["log_n"] == 20)
    seed = int(c["seed"], 16)
    tabs = [cref.gen_table(0, seed, k, c["log_n"]) for k in range(c["m"])]
    claim = cref.product_sum(0, tabs, c["log_n"])
    assert [hex(int(x)) for x in claim] == c["claim_mont_limbs"]
    rp, ch, fin = cref.prove(0, tabs, c["log_n"], c["degree"], claim, False, fast=True)
    assert cref.keccak256(rp.tobytes() + ch.tobytes()).hex() == c["proof_keccak"]
    assert cref.keccak256(fin.tobytes()).hex() == c["finals_keccak"]
    # the streamlined prover used for the digests is the reference-shaped one, bit for bit (small size)
    t2 = [cref.gen_table(0, seed, k, 10) for k in range(3)]
    cl = cref.product_sum(0, t2, 10)
    a = cref.prove(0, t2, 10, 3, cl, False, fast=True)
    b = cref.prove(0, t2, 10, 3, cl, False, fast=False)
    assert all((x == y).all() for x, y in zip(a, b))


def test_keccak_sponge_pinned_by_hashlib_sha3_256():
    """An implementation this repo did not write: Python's hashlib has no Keccak-256, but its SHA3-256 is the same
    Keccak-f[1600], rate 136 and capacity 512 — only the domain byte of the padding differs (0x06 instead of 0x01).
    Running the oracle's sponge with 0x06 must therefore reproduce hashlib.sha3_256 on every length around the block
    boundaries and on multi-block messages in uneven chunks; the 0x01 byte itself is pinned by the published
    Keccak-256 digests of "" and "abc" (test_keccak_kats)."""
    import hashlib

    msg = bytes((i * 131 + 7) & 0xFF for i in range(1000))
    for n in list(range(0, 300)) + [407, 408, 409, 543, 544, 545, 999, 1000]:
        h = O.Keccak256(domain=0x06)
        h.update(msg[:n])
        assert h.finalize_reset() == hashlib.sha3_256(msg[:n]).digest(), n
    h = O.Keccak256(domain=0x06)
    for a, b in [(0, 1), (1, 137), (137, 138), (138, 600), (600, 1000)]:
        h.update(msg[a:b])
    assert h.finalize_reset() == hashlib.sha3_256(msg).digest()
    # finalize_reset really resets
    h.update(b"abc")
    assert h.finalize_reset() == hashlib.sha3_256(b"abc").digest()
    # and the reference's variant (0x01) on the published vectors
    assert O.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert O.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"


def test_field_constants_follow_from_the_curves_public_parameters():
    """The moduli are not free constants: both scalar fields are r = z^4 - z^2 + 1 of the BLS12 family parameter
    (BLS12-381: z = -0xd201000000010000, BLS12-377: z = 0x8508c00000000001).  The multiplicative generators ark-ff
    declares (7, 22) must be quadratic non-residues for g^((r-1)/2^s) to be a PRIMITIVE 2^s-th root of unity, and the
    2^32-th root derived from 7 is the `ROOT_OF_UNITY` the zkcrypto `bls12_381::Scalar` type publishes for the same
    generator — a second library agreeing on the NTT's root choice, which the reference's only fft test (a round trip,
    fft/src/lib.rs:78-82) does not pin."""
    z381, z377 = -0xD201000000010000, 0x8508C00000000001
    assert z381**4 - z381**2 + 1 == O.BLS12_381_FR.p
    assert z377**4 - z377**2 + 1 == O.BLS12_377_FR.p
    for Fq in (O.BLS12_381_FR, O.BLS12_377_FR):
        assert pow(Fq.generator, (Fq.p - 1) // 2, Fq.p) == Fq.p - 1
        w = Fq.get_root_of_unity(1 << Fq.two_adicity)
        assert pow(w, 1 << (Fq.two_adicity - 1), Fq.p) == Fq.p - 1  # exact order 2^s
        assert Fq.get_root_of_unity(1 << (Fq.two_adicity + 1)) is None
    assert (O.BLS12_381_FR.two_adicity, O.BLS12_377_FR.two_adicity) == (32, 47)
    assert O.BLS12_381_FR.get_root_of_unity(1 << 32) == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B


def test_multithreaded_streamlined_prover_is_bit_identical(cref):
    """oracle/cpu_ref.c::zko_prove_fast_mt (bench.py's `streamlined_all_cores` CPU figure): same proof as the
    single-threaded provers, whatever the thread count (modular sums are order independent)."""
    for fid, n, m, d in [(0, 14, 3, 3), (1, 13, 2, 2), (0, 12, 1, 1)]:
        tabs = [cref.gen_table(fid, 9, k, n) for k in range(m)]
        claim = cref.product_sum(fid, tabs, n)
        ref = cref.prove(fid, tabs, n, d, claim, False)
        for threads in (1, 3, 8):
            got = cref.prove(fid, tabs, n, d, claim, False, fast=True, threads=threads)
            assert all((x == y).all() for x, y in zip(ref, got)), (fid, n, m, d, threads)


def test_streaming_generated_prover_is_bit_identical(cref):
    """oracle/cpu_ref.c::zko_prove_generated_mt (the offline oracle behind the 2^27 .. 2^30 digests of
    tests/golden/fullsize_digests.json: tables regenerated from the seed instead of held at full size): same claim and
    proof as the reference-shaped prover on the materialised tables."""
    for fid, n, m, d, threads in [(0, 13, 3, 3, 1), (0, 14, 3, 3, 5), (1, 12, 2, 2, 8), (0, 11, 1, 1, 2), (0, 1, 3, 3, 4), (0, 2, 2, 3, 1)]:
        tabs = [cref.gen_table(fid, 0x5EED000000000001, k, n) for k in range(m)]
        claim = cref.product_sum(fid, tabs, n)
        ref = cref.prove(fid, tabs, n, d, claim, False)
        gclaim, *got = cref.prove_generated(fid, 0x5EED000000000001, m, n, d, threads)
        assert (gclaim == claim).all(), (fid, n, m, d)
        assert all((x == y).all() for x, y in zip(ref, got)), (fid, n, m, d, threads)
