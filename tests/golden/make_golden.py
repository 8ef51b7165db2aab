#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the pure-Python big-int oracle (oracle/zkoracle.py).

Provenance: the Rust reference cannot run here, so these vectors are produced by the line-by-line
restatement; the survey's independently derived vectors are embedded as
`survey_*` entries and must agree (asserted below).  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import zkoracle as O  # noqa: E402

F = O.BLS12_381_FR
hx = lambda v: hex(v)


def prove_case(tables, degree, claim, absorb, name, field=F):
    pp = O.ProductPoly([O.MultiLinearPolynomial(field, len(tables[0]).bit_length() - 1, t) for t in tables])
    pr = O.SumcheckProver(degree)
    if absorb:
        tr = O.Transcript()
        tr.append(pp.to_bytes())
        proof, ch = pr.prove_internal(pp, claim, tr)
    else:
        proof, ch = pr.prove_partial(pp, claim)
    finals = [q.evaluations[0] for q in pr.final_poly.polynomials]
    return {
        "name": name, "field": field.field_id, "degree": degree, "absorb": absorb, "claim": hx(claim),
        "round_polys": [[hx(x) for x in rp] for rp in proof.round_polys], "challenges": [hx(c) for c in ch],
        "finals": [hx(x) for x in finals],
    }


def main():
    out = {}
    # --- reference-test fixtures (sumcheck/src/lib.rs:53-122) ---
    t_2ab3bc = [0, 0, 0, 3, 0, 0, 2, 5]
    cases = []
    c = prove_case([t_2ab3bc], 1, 10, True, "B1_prove_2ab3bc_sum10"); c["tables"] = [t_2ab3bc]; cases.append(c)
    assert c["challenges"][0] == "0x3ffa109ae4d4a5ca68127a16b99d9e8a23103f8c812c11911756104955b38be2"
    assert c["finals"][0] == "0x24afe72dcaf12414ed96c5f8c84897d74c8aaf8dad208059e4b37399b5185478"
    c = prove_case([t_2ab3bc], 1, 10, False, "B2_prove_partial_2ab3bc_sum10"); c["tables"] = [t_2ab3bc]; cases.append(c)
    assert c["challenges"][2] == "0x66409d0fc3dfc9f8e535d27c8e5d32cf467b235c4170f05b3dc454f605220bda"
    c = prove_case([[3, 3, 5, 5], [0, 0, 0, 1]], 2, 5, True, "B3_prove_deg2_sum5"); c["tables"] = [[3, 3, 5, 5], [0, 0, 0, 1]]; cases.append(c)
    assert c["challenges"][1] == "0x67d991dbcc7e7384d27b32213c70ef2fe12a6f8b970cd34d5c8265470a656dd9"
    c = prove_case([t_2ab3bc], 1, 12, True, "B4_prove_2ab3bc_wrong_sum12"); c["tables"] = [t_2ab3bc]; cases.append(c)
    out["small_cases"] = cases

    # --- seeded synthetic tables (SURVEY 8d generator) ---
    seeded = []
    for (n, m, d) in [(4, 3, 3), (10, 3, 3), (10, 2, 2), (10, 1, 1), (7, 2, 3), (7, 3, 1), (6, 4, 4), (5, 1, 3), (1, 3, 3), (2, 2, 2)]:
        tabs = [O.gen_table(F, O.DEFAULT_SEED, k, n) for k in range(m)]
        pp = O.ProductPoly([O.MultiLinearPolynomial(F, n, t) for t in tabs])
        claim = sum(pp.prod_reduce()) % F.p
        for absorb in (False, True):
            c = prove_case(tabs, d, claim, absorb, f"seeded_n{n}_m{m}_d{d}_{'prove' if absorb else 'partial'}")
            c.update({"n": n, "m": m, "seed": hx(O.DEFAULT_SEED)})
            seeded.append(c)
    out["seeded_cases"] = seeded
    s = {c["name"]: c for c in seeded}
    assert s["seeded_n4_m3_d3_partial"]["challenges"][-1] == "0x101bc4a4806901f5b1b1a21e732dc00f6d3b83e2da47d2f939cb0ab60946552a"
    assert s["seeded_n10_m3_d3_partial"]["challenges"][-1] == "0x12655342632ba61485296855ffd521360a9b8ac0dac7a184d0633da0f3d5ec49"
    assert s["seeded_n10_m2_d2_partial"]["challenges"][-1] == "0x3292aae7d282bed9a1d6b65b14989dde32c00f954d76df1e070d151fe92be34d"
    assert s["seeded_n10_m1_d1_partial"]["challenges"][-1] == "0x6cf60ee199525e08ad77b88f819535198230ff49d5c77ceafef38743cbe171e9"

    # --- generator anchors ---
    out["generator"] = {
        "seed": hx(O.DEFAULT_SEED),
        "elem_k0_i0": hx(O.gen_element(O.DEFAULT_SEED, 0, 0)),
        "elem_k0_i0_mont": hx(F.to_mont(O.gen_element(O.DEFAULT_SEED, 0, 0))),
        "elem_k0_i1": hx(O.gen_element(O.DEFAULT_SEED, 0, 1)),
        "elem_k2_i2p30m1": hx(O.gen_element(O.DEFAULT_SEED, 2, 2**30 - 1)),
        "elem377_k1_i5": hx(O.gen_element(O.DEFAULT_SEED, 1, 5) % O.BLS12_377_FR.p),
    }
    assert out["generator"]["elem_k0_i0_mont"] == "0x4965a082553a274fae5c777c101165db62c3e25d617068994e92cce83d7927d1"

    # --- transcript ---
    tr = O.Transcript(); tr.append(b"zk-b200 transcript golden"); tr.append(bytes(range(200)))
    ch = [hx(tr.sample_field_element(F)) for _ in range(3)]
    tr.append(b"more"); ch.append(hx(tr.sample_field_element(O.BLS12_377_FR)))
    out["transcript"] = {"keccak_empty": O.keccak256(b"").hex(), "keccak_abc": O.keccak256(b"abc").hex(),
                         "keccak_200x61": O.keccak256(b"a" * 200).hex(), "keccak_136x00": O.keccak256(bytes(136)).hex(),
                         "challenges": ch}
    assert out["transcript"]["keccak_empty"] == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert out["transcript"]["keccak_abc"] == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"

    # --- partial_evaluate with general initial_var / multi-assignment ---
    pe = []
    n = 6
    tab = O.gen_table(F, 11, 0, n)
    P = O.MultiLinearPolynomial(F, n, tab)
    for iv, assigns in [(0, [5]), (0, [0]), (0, [1]), (5, [7]), (2, [3, 4]), (1, [2, 3, 4, 5]), (0, [9, 8, 7, 6, 5, 4]), (3, [0, 1, 2])]:
        pe.append({"n": n, "seed": 11, "initial_var": iv, "assignments": assigns,
                   "out": [hx(x) for x in P.partial_evaluate(iv, assigns).evaluations]})
    out["partial_evaluate"] = pe

    # --- FFT goldens (forward values are NOT pinned by the reference; cross-checked with the naive DFT) ---
    ffts = []
    for FF in (O.BLS12_381_FR, O.BLS12_377_FR):
        for log_n in (0, 1, 2, 3, 6):
            vals = [O.gen_element(3 + log_n, FF.field_id, i) % FF.p for i in range(1 << log_n)]
            fw = O.fft(FF, vals)
            assert fw == O.naive_dft(FF, vals, FF.get_root_of_unity(1 << log_n))
            assert O.ifft(FF, fw) == vals
            ffts.append({"field": FF.field_id, "log_n": log_n, "seed": 3 + log_n, "in": [hx(v) for v in vals], "fft": [hx(v) for v in fw]})
    f377 = O.fft(O.BLS12_377_FR, [0, 2, 34, 3434])
    assert hx(f377[1]) == "0x12ab655e9a2ca374be69d55d4993c524e67d4bbf4fffeb1a92277ffffffff277"
    ffts.append({"field": 1, "log_n": 2, "seed": None, "in": [hx(v) for v in [0, 2, 34, 3434]], "fft": [hx(v) for v in f377]})
    out["fft"] = ffts
    out["roots"] = {"w4_381": hx(O.BLS12_381_FR.get_root_of_unity(4)), "w4_377": hx(O.BLS12_377_FR.get_root_of_unity(4)),
                    "w2p32_381": hx(O.BLS12_381_FR.get_root_of_unity(1 << 32)), "w2p47_377": hx(O.BLS12_377_FR.get_root_of_unity(1 << 47))}
    assert out["roots"]["w2p32_381"] == hx(10238227357739495823651030575849232062558860180284477541189508159991286009131)
    assert out["roots"]["w2p47_377"] == hx(8065159656716812877374967518403273466521432693661810619979959746626482506078)

    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "vectors.json"))


if __name__ == "__main__":
    main()
