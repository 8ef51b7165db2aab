"""CPU oracle (pure-Python big-int restatement) of the iammadab/zk sumcheck / MLE / FFT path.

TEST INFRASTRUCTURE ONLY.  Nothing under `zk_b200/` (the product) may import this module;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs use `oracle/`, and there only as the checker.

The Rust reference cannot be built in this environment (no rustc/cargo, arithmetic lives in
un-vendored crates ark-ff 0.5.0 / ark-bls12-381 0.5.0 / ark-bls12-377 0.5.0 / sha3 0.10.8), so
this file restates the reference line by line over Python integers.  Each function cites the
reference file:line (relative to /root/reference) it follows.  Pinning: every known-answer
test the reference holds for this path is reproduced in tests/test_oracle_kats.py; values the
reference never asserts (transcript bytes, round polynomials, forward NTT values) are pinned
only against the published Keccak-256 KATs, a naive DFT and SURVEY.md Appendix B.
"""
from __future__ import annotations

# --------------------------------------------------------------------------------------
# Fields (third-party: ark-bls12-381 0.5.0 / ark-bls12-377 0.5.0 `Fr`; ark-ff Fp64 mod 17 in tests)
# --------------------------------------------------------------------------------------


class Field:
    """Prime field description: modulus, multiplicative generator, two-adicity (ark-ff FftField)."""

    def __init__(self, name: str, p: int, generator: int, field_id: int):
        self.name = name
        self.p = p
        self.generator = generator
        self.field_id = field_id
        s, t = 0, p - 1
        while t % 2 == 0:
            s, t = s + 1, t // 2
        self.two_adicity = s
        self.R = (1 << 256) % p  # Montgomery radix for 4x64 limbs (ark-ff MontBackend<_, 4>)
        self.R2 = (self.R * self.R) % p
        self.Rinv = pow(self.R, -1, p)

    # ark-ff `F::from(u64)`, `from_be_bytes_mod_order`, `into_bigint().to_bytes_be()`
    def from_be_bytes_mod_order(self, b: bytes) -> int:
        return int.from_bytes(b, "big") % self.p

    def to_bytes_be(self, x: int) -> bytes:
        return (x % self.p).to_bytes(32, "big")

    def inv(self, x: int) -> int:
        return pow(x, -1, self.p)

    # ark-ff `FftField::get_root_of_unity(n)`: g^((p-1)/n) for n a power of two <= 2^two_adicity
    def get_root_of_unity(self, n: int):
        if n == 0 or n & (n - 1) or n > (1 << self.two_adicity):
            return None
        return pow(self.generator, (self.p - 1) // n, self.p)

    def to_mont(self, x: int) -> int:
        return (x * self.R) % self.p

    def from_mont(self, x: int) -> int:
        return (x * self.Rinv) % self.p

    def mont_limbs(self, x: int):
        """Canonical int -> the 4 little-endian u64 Montgomery limbs arkworks keeps in memory."""
        m = self.to_mont(x % self.p)
        return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]

    def from_mont_limbs(self, limbs) -> int:
        m = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
        return self.from_mont(m)


BLS12_381_FR = Field(
    "bls12_381_fr",
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    7,
    0,
)
BLS12_377_FR = Field(
    "bls12_377_fr",
    0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    22,
    1,
)
F17 = Field("f17", 17, 3, 99)  # Fp64 mod 17, generator 3 (polynomial/src/univariate_poly.rs:237-241)
FIELDS = {0: BLS12_381_FR, 1: BLS12_377_FR}

# --------------------------------------------------------------------------------------
# Keccak-256 (sha3 0.10.8 `Keccak256`: Keccak[r=1088,c=512], pad10*1 with domain byte 0x01)
# --------------------------------------------------------------------------------------

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f1600(A):
    """Keccak-f[1600] on a 5x5 list-of-lists state A[x][y] (FIPS 202 section 3.2-3.3)."""
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    return A


class Keccak256:
    """Incremental Keccak-256 with the `digest::Digest` surface the reference uses
    (`update`, `finalize_reset`): transcript/src/lib.rs:2,17,22."""

    RATE = 136
    DOMAIN = 0x01  # original Keccak padding; 0x06 would be SHA3-256 (same permutation, rate and capacity)

    def __init__(self, domain: int = 0x01):
        self.DOMAIN = domain
        self._reset()

    def _reset(self):
        self.state = [[0] * 5 for _ in range(5)]
        self.buf = b""

    def _absorb_block(self, block: bytes):
        for i in range(self.RATE // 8):
            lane = int.from_bytes(block[8 * i : 8 * i + 8], "little")
            self.state[i % 5][i // 5] ^= lane
        self.state = keccak_f1600(self.state)

    def update(self, data: bytes):
        self.buf += bytes(data)
        off = 0
        while len(self.buf) - off >= self.RATE:
            self._absorb_block(self.buf[off : off + self.RATE])
            off += self.RATE
        self.buf = self.buf[off:]

    def finalize_reset(self) -> bytes:
        pad = bytearray(self.buf) + bytearray(self.RATE - len(self.buf))
        pad[len(self.buf)] ^= self.DOMAIN  # 0x01: original Keccak domain/padding byte (SHA3 would be 0x06)
        pad[self.RATE - 1] ^= 0x80
        self._absorb_block(bytes(pad))
        out = b"".join(self.state[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
        self._reset()
        return out


def keccak256(data: bytes) -> bytes:
    h = Keccak256()
    h.update(data)
    return h.finalize_reset()


class Transcript:
    """transcript/src/lib.rs:5-35."""

    def __init__(self):  # :10-14
        self.hasher = Keccak256()

    def append(self, new_data: bytes):  # :16-18
        self.hasher.update(new_data)

    def _sample_challenge(self) -> bytes:  # :20-25  finalize_reset, then re-absorb the digest
        result_hash = self.hasher.finalize_reset()
        self.hasher.update(result_hash)
        return result_hash

    def sample_field_element(self, F: Field) -> int:  # :27-30
        return F.from_be_bytes_mod_order(self._sample_challenge())

    def sample_n_field_elements(self, F: Field, n: int):  # :32-34
        return [self.sample_field_element(F) for _ in range(n)]


# --------------------------------------------------------------------------------------
# polynomial::multilinear::pairing_index  (polynomial/src/multilinear/pairing_index.rs)
# --------------------------------------------------------------------------------------


def mask(n: int) -> int:  # :24-26
    return (1 << n) - 1


def insert_bit(val: int, index: int, bit: int) -> int:  # :17-21
    high = val >> index
    low = val & mask(index)
    return high << (index + 1) | bit << index | low


def index_pair(n_vars: int, index: int):  # :2-9
    base_no_of_vars = n_vars - 1
    if base_no_of_vars < 0 or base_no_of_vars - index < 0:
        raise OverflowError("attempt to subtract with overflow")  # Rust debug-build panic
    no_of_pairs = 1 << base_no_of_vars
    for val in range(no_of_pairs):
        insert_0 = insert_bit(val, base_no_of_vars - index, 0)
        yield insert_0, insert_0 | (1 << (base_no_of_vars - index))


# --------------------------------------------------------------------------------------
# polynomial::multilinear::evaluation_form  (polynomial/src/multilinear/evaluation_form.rs)
# --------------------------------------------------------------------------------------


class OracleError(Exception):
    """Carries the reference's `&'static str` error message."""


class MultiLinearPolynomial:
    def __init__(self, F: Field, n_vars: int, evaluations):  # new :15-27
        if len(evaluations) != (1 << n_vars):
            raise OracleError("evaluation vec len should equal 2^n_vars")
        self.F = F
        self._n_vars = n_vars
        self.evaluations = [e % F.p for e in evaluations]

    def n_vars(self) -> int:  # :30-32
        return self._n_vars

    def partial_evaluate(self, initial_var: int, assignments):  # :40-80
        p = self.F.p
        new_evaluations = list(self.evaluations)  # :49 clone
        for i, assignment in enumerate(assignments):  # :54
            a = assignment % p
            for k, (left_pos, right_pos) in enumerate(index_pair(self._n_vars - i, initial_var)):  # :55-56
                left = new_evaluations[left_pos]
                right = new_evaluations[right_pos]
                if a == 0:  # :61
                    new_evaluations[k] = left
                elif a == 1:  # :62
                    new_evaluations[k] = right
                else:  # :68  left - r (left - right)
                    new_evaluations[k] = (left - a * (left - right)) % p
        new_n_vars = self._n_vars - len(assignments)  # :75
        if new_n_vars < 0:
            raise OverflowError("attempt to subtract with overflow")
        return MultiLinearPolynomial(self.F, new_n_vars, new_evaluations[: 1 << new_n_vars])  # :76-79

    def evaluate(self, assignments) -> int:  # :83-89
        if len(assignments) != self._n_vars:
            raise OracleError("evaluate must assign to all variables")
        return self.partial_evaluate(0, assignments).evaluations[0]

    def evaluation_slice(self):  # :92-94
        return self.evaluations

    def to_bytes(self) -> bytes:  # :97-103
        return b"".join(self.F.to_bytes_be(e) for e in self.evaluations)

    def __eq__(self, other):
        return self._n_vars == other._n_vars and self.evaluations == other.evaluations


# --------------------------------------------------------------------------------------
# polynomial::product_poly  (polynomial/src/product_poly.rs)
# --------------------------------------------------------------------------------------


class ProductPoly:
    def __init__(self, polynomials):  # new :14-32
        if len(polynomials) == 0:
            raise OracleError("cannot create product polynomial from empty polynomials")
        expected = polynomials[0].n_vars()
        if not all(q.n_vars() == expected for q in polynomials):
            raise OracleError(
                "cannot create product polynomial from polynomial that don't share the same number of variables"
            )
        self._n_vars = expected
        self.polynomials = list(polynomials)
        self.F = polynomials[0].F

    def evaluate(self, assignments) -> int:  # :36-44
        if len(assignments) != self._n_vars:
            raise OracleError("evaluate must assign to all variables")
        product = 1
        for q in self.polynomials:
            product = (product * q.evaluate(assignments)) % self.F.p
        return product

    def partial_evaluate(self, initial_var: int, assignments):  # :48-63
        return ProductPoly([q.partial_evaluate(initial_var, assignments) for q in self.polynomials])

    def prod_reduce(self):  # :66-74
        result = list(self.polynomials[0].evaluation_slice())
        for q in self.polynomials[1:]:
            for i, e in enumerate(q.evaluation_slice()):
                result[i] = (result[i] * e) % self.F.p
        return result

    def to_bytes(self) -> bytes:  # :77-83  factor-major, index-minor
        return b"".join(q.to_bytes() for q in self.polynomials)

    def n_vars(self) -> int:  # :86-88
        return self._n_vars

    def clone(self):
        return ProductPoly([MultiLinearPolynomial(q.F, q.n_vars(), list(q.evaluations)) for q in self.polynomials])


class SumOfProductsPoly:
    """P(x) = sum_t prod_{k in terms[t]} polynomials[k](x)   (SURVEY.md 8f-4).

    NOT in the reference: its ProductPoly is a single product (product_poly.rs:4-10) and the GKR crate readme.md:9
    links is absent.  This class is the smallest extension that lets the reference's own prover loop
    (SumcheckProver.prove_internal, prover.rs:33-73) run unchanged over a GKR layer polynomial
    add.Wb + add.Wc + mul.Wb.Wc: it offers the same four methods the loop calls on a ProductPoly — n_vars,
    partial_evaluate, prod_reduce (here: the table of P, sum of the terms' element-wise products) and to_bytes
    (the tables in order).  With a single term listing every table once it IS ProductPoly (tests pin that)."""

    def __init__(self, polynomials, terms):
        if len(polynomials) == 0 or len(terms) == 0 or any(len(t) == 0 for t in terms):
            raise OracleError("cannot create product polynomial from empty polynomials")
        expected = polynomials[0].n_vars()
        if not all(q.n_vars() == expected for q in polynomials):
            raise OracleError(
                "cannot create product polynomial from polynomial that don't share the same number of variables"
            )
        if any(k < 0 or k >= len(polynomials) for t in terms for k in t):
            raise OracleError("invalid argument")
        self._n_vars = expected
        self.polynomials = list(polynomials)
        self.terms = [list(t) for t in terms]
        self.F = polynomials[0].F

    def n_vars(self) -> int:
        return self._n_vars

    def partial_evaluate(self, initial_var: int, assignments):
        return SumOfProductsPoly([q.partial_evaluate(initial_var, assignments) for q in self.polynomials], self.terms)

    def prod_reduce(self):
        p = self.F.p
        result = [0] * (1 << self._n_vars)
        for term in self.terms:
            prod = list(self.polynomials[term[0]].evaluation_slice())
            for k in term[1:]:
                for i, e in enumerate(self.polynomials[k].evaluation_slice()):
                    prod[i] = (prod[i] * e) % p
            for i, e in enumerate(prod):
                result[i] = (result[i] + e) % p
        return result

    def evaluate(self, assignments) -> int:
        if len(assignments) != self._n_vars:
            raise OracleError("evaluate must assign to all variables")
        vals = [q.evaluate(assignments) for q in self.polynomials]
        return self.combine(vals)

    def combine(self, table_values) -> int:
        total = 0
        for term in self.terms:
            prod = 1
            for k in term:
                prod = (prod * table_values[k]) % self.F.p
            total = (total + prod) % self.F.p
        return total

    def to_bytes(self) -> bytes:
        return b"".join(q.to_bytes() for q in self.polynomials)

    def clone(self):
        return SumOfProductsPoly(
            [MultiLinearPolynomial(q.F, q.n_vars(), list(q.evaluations)) for q in self.polynomials], self.terms
        )


# --------------------------------------------------------------------------------------
# polynomial::univariate_poly  (verifier side only; polynomial/src/univariate_poly.rs)
# --------------------------------------------------------------------------------------


class UnivariatePolynomial:
    def __init__(self, F: Field, coefficients):  # :16-18
        self.F = F
        self.coefficients = [c % F.p for c in coefficients]

    def evaluate(self, x: int) -> int:  # :29-40  Horner
        acc = 0
        for c in reversed(self.coefficients):
            acc = (acc * x + c) % self.F.p
        return acc

    @classmethod
    def interpolate(cls, F: Field, ys):  # :43-49
        return cls.interpolate_xy(F, list(range(len(ys))), ys)

    @classmethod
    def interpolate_xy(cls, F: Field, xs, ys):  # :54-80
        p = F.p
        result = cls(F, [])
        for li, (x, y) in enumerate(zip(xs, ys)):
            basis = cls(F, [1])
            for xi, xv in enumerate(xs):
                if xi == li:
                    continue
                numerator = cls(F, [(-xv) % p, 1])
                denominator = F.inv((x - xv) % p)
                basis = basis.mul(numerator.mul(cls(F, [denominator])))
            result = result.add(basis.mul(cls(F, [y])))
        return result

    def is_zero(self):  # :83-85
        return len(self.coefficients) == 0

    def degree(self):  # :88-94
        return 0 if not self.coefficients else len(self.coefficients) - 1

    def add(self, other):  # :157-183
        if self.is_zero():
            return UnivariatePolynomial(self.F, other.coefficients)
        if other.is_zero():
            return UnivariatePolynomial(self.F, self.coefficients)
        if len(self.coefficients) >= len(other.coefficients):
            new, oth = list(self.coefficients), other.coefficients
        else:
            new, oth = list(other.coefficients), self.coefficients
        for i in range(len(oth)):
            new[i] = (new[i] + oth[i]) % self.F.p
        return UnivariatePolynomial(self.F, new)

    def mul(self, other):  # :185-209
        if self.is_zero() or other.is_zero():
            return UnivariatePolynomial(self.F, [])
        out = [0] * (self.degree() + other.degree() + 1)
        for i in range(self.degree() + 1):
            for j in range(other.degree() + 1):
                out[i + j] = (out[i + j] + self.coefficients[i] * other.coefficients[j]) % self.F.p
        return UnivariatePolynomial(self.F, out)


# --------------------------------------------------------------------------------------
# sumcheck  (sumcheck/src/lib.rs, prover.rs, verifier.rs)
# --------------------------------------------------------------------------------------


class SumcheckProof:  # sumcheck/src/lib.rs:8-11
    def __init__(self, sum_, round_polys):
        self.sum = sum_
        self.round_polys = round_polys


class SubClaim:  # sumcheck/src/lib.rs:17-20
    def __init__(self, sum_, challenges):
        self.sum = sum_
        self.challenges = challenges


def field_elements_to_bytes(F: Field, elems) -> bytes:  # sumcheck/src/lib.rs:23-29
    return b"".join(F.to_bytes_be(e) for e in elems)


class SumcheckProver:
    """sumcheck/src/prover.rs; `max_var_degree` is the const generic MAX_VAR_DEGREE."""

    def __init__(self, max_var_degree: int):
        self.D = max_var_degree

    def prove(self, poly: ProductPoly, sum_: int):  # :15-20
        transcript = Transcript()
        transcript.append(poly.to_bytes())
        return self.prove_internal(poly, sum_, transcript)[0]

    def prove_partial(self, poly: ProductPoly, sum_: int):  # :24-30
        transcript = Transcript()
        return self.prove_internal(poly, sum_, transcript)

    def prove_internal(self, poly: ProductPoly, sum_: int, transcript: Transcript):  # :33-73
        F = poly.F
        round_polys, challenges = [], []
        transcript.append(F.to_bytes_be(sum_))  # :42
        for _ in range(poly.n_vars()):  # :44
            round_poly = []
            for i in range(self.D + 1):  # :49
                vals = poly.partial_evaluate(0, [i]).prod_reduce()  # :51-52
                round_poly.append(sum(vals) % F.p)  # :53-54
            transcript.append(field_elements_to_bytes(F, round_poly))  # :59
            challenge = transcript.sample_field_element(F)  # :62
            poly = poly.partial_evaluate(0, [challenge])  # :64
            round_polys.append(round_poly)
            challenges.append(challenge)
        self.final_poly = poly  # (oracle extra: the fully folded factors, for parity dumps)
        return SumcheckProof(sum_ % F.p, round_polys), challenges


class SumcheckVerifier:
    """sumcheck/src/verifier.rs."""

    @staticmethod
    def verify(poly: ProductPoly, proof: SumcheckProof) -> bool:  # :15-33
        if len(proof.round_polys) != poly.n_vars():
            raise OracleError("invalid proof: require 1 round poly for each variable in poly")
        transcript = Transcript()
        transcript.append(poly.to_bytes())
        subclaim = SumcheckVerifier.verify_internal(poly.F, proof, transcript)
        try:
            initial_poly_eval = poly.evaluate(subclaim.challenges)
        except OracleError:
            raise OracleError("couldn't evaluate initial poly")
        return initial_poly_eval == subclaim.sum

    @staticmethod
    def verify_partial(F: Field, proof: SumcheckProof) -> SubClaim:  # :38-41
        return SumcheckVerifier.verify_internal(F, proof, Transcript())

    @staticmethod
    def verify_internal(F: Field, proof: SumcheckProof, transcript: Transcript) -> SubClaim:  # :44-78
        challenges = []
        transcript.append(F.to_bytes_be(proof.sum))  # :50
        claimed_sum = proof.sum % F.p
        for round_poly in proof.round_polys:  # :54
            transcript.append(field_elements_to_bytes(F, round_poly))  # :56
            u = UnivariatePolynomial.interpolate(F, round_poly)  # :58
            p_0, p_1 = u.evaluate(0), u.evaluate(1)  # :61-62
            if claimed_sum != (p_0 + p_1) % F.p:  # :64-66
                raise OracleError("verifier check failed: claimed_sum != p(0) + p(1)")
            challenge = transcript.sample_field_element(F)  # :69
            claimed_sum = u.evaluate(challenge)  # :70
            challenges.append(challenge)
        return SubClaim(claimed_sum, challenges)


# --------------------------------------------------------------------------------------
# fft  (fft/src/lib.rs)
# --------------------------------------------------------------------------------------


def split_even_odd(data):  # :48-61
    return data[0::2], data[1::2]


def fft_internal(F: Field, values, omega: int):  # :21-46
    if len(values) == 1:
        return list(values)
    n = len(values)
    if n & (n - 1):
        raise ValueError("values must be a power of 2")  # panic :29
    p = F.p
    even, odd = split_even_odd(values)
    w2 = (omega * omega) % p
    even_evals = fft_internal(F, even, w2)
    odd_evals = fft_internal(F, odd, w2)
    out = [0] * n
    for i in range(n // 2):  # :40-43
        out[i] = (even_evals[i] + pow(omega, i, p) * odd_evals[i]) % p
        out[i + n // 2] = (even_evals[i] + pow(omega, i + n // 2, p) * odd_evals[i]) % p
    return out


def fft(F: Field, coefficients):  # :4-8
    omega = F.get_root_of_unity(len(coefficients))
    if omega is None:
        raise ValueError("called `Option::unwrap()` on a `None` value")
    return fft_internal(F, coefficients, omega)


def ifft(F: Field, evaluations):  # :11-19
    n = len(evaluations)
    omega = F.get_root_of_unity(n)
    if omega is None:
        raise ValueError("called `Option::unwrap()` on a `None` value")
    omega = F.inv(omega)
    n_inv = F.inv(n % F.p)
    return [(v * n_inv) % F.p for v in fft_internal(F, evaluations, omega)]


def naive_dft(F: Field, values, omega: int):
    """O(n^2) definition X[i] = sum_j a_j w^(ij): the cross-check for forward-transform values,
    which the reference never pins (fft/src/lib.rs:78-82 is a round trip only)."""
    n, p = len(values), F.p
    return [sum(values[j] * pow(omega, (i * j) % n, p) for j in range(n)) % p for i in range(n)]


# --------------------------------------------------------------------------------------
# Deterministic synthetic tables: counter-based splitmix64, keyed by GLOBAL index
# --------------------------------------------------------------------------------------

DEFAULT_SEED = 0x5EED000000000001
_GOLDEN = 0x9E3779B97F4A7C15


def splitmix64_mix(z: int) -> int:
    z &= _M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z


def gen_element(seed: int, table_id: int, index: int) -> int:
    """Canonical value v < 2^254 of entry `index` of synthetic table `table_id`."""
    limbs = []
    for l in range(4):
        ctr = (((table_id << 40) + index) * 4 + l) & _M64
        limbs.append(splitmix64_mix((seed + _GOLDEN * (ctr + 1)) & _M64))
    limbs[3] &= 0x3FFFFFFFFFFFFFFF
    return sum(l << (64 * i) for i, l in enumerate(limbs))


def gen_table(F: Field, seed: int, table_id: int, n_vars: int):
    return [gen_element(seed, table_id, i) % F.p for i in range(1 << n_vars)]
