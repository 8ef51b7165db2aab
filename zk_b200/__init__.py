"""zk_b200 — host-side mirror of the iammadab/zk `polynomial` / `sumcheck` / `transcript` / `fft` crate
APIs over the B200-native C ABI (include/zk_b200.h, libzk_b200.so).

Same names, argument meaning and error strings as the reference (file:line cited per item), so the
parity tests read like the reference's own tests.  Field elements are Python ints (canonical values,
like `Fr::from(..)`); tables live on the GPU.  All arithmetic happens in the CUDA library — there is no
Python or CPU fallback, and construction fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import lib

BLS12_381_FR = 0  # ark_bls12_381::Fr
BLS12_377_FR = 1  # ark_bls12_377::Fr
MODULUS = {
    BLS12_381_FR: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    BLS12_377_FR: 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
}
DEFAULT_SEED = 0x5EED000000000001


class ZkError(Exception):
    """An `Err(&'static str)` (or panic message) of the reference; `.message` is the literal string."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status
        self.message = message


def _check(status: int, ctx=None):
    if status != _ffi.OK:
        msg = lib().zk_status_string(status).decode()
        raise ZkError(status, msg)


# ---- element conversions (ark-ff From/into_bigint) -------------------------------------------------
def _limbs_of(vals: Sequence[int]) -> np.ndarray:
    raw = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(raw, dtype="<u8").reshape(-1, 4).copy()


def to_mont(field: int, vals: Sequence[int]) -> np.ndarray:
    """Python ints -> (n,4) uint64 Montgomery limbs (values are first reduced mod p, like `Fr::from`)."""
    p = MODULUS[field]
    canon = _limbs_of([int(v) % p for v in vals])
    out = np.empty_like(canon)
    _check(lib().zk_field_from_canonical(field, canon.ctypes.data, out.ctypes.data, canon.shape[0]))
    return out


def from_mont(field: int, mont: np.ndarray) -> List[int]:
    mont = np.ascontiguousarray(mont, dtype=np.uint64).reshape(-1, 4)
    canon = np.empty_like(mont)
    _check(lib().zk_field_to_canonical(field, mont.ctypes.data, canon.ctypes.data, mont.shape[0]))
    raw = canon.tobytes()
    return [int.from_bytes(raw[32 * i : 32 * i + 32], "little") for i in range(mont.shape[0])]


class Context:
    """One GPU.  `Context.default()` is the process-wide single-GPU context used by the mirrored API."""

    _default: Optional["Context"] = None

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_id: Optional[bytes] = None):
        h = C.c_void_p()
        if world > 1:
            buf = C.create_string_buffer(nccl_id, 128)
            _check(lib().zk_ctx_create_sharded(device, rank, world, C.cast(buf, C.c_void_p), C.byref(h)))
        else:
            _check(lib().zk_ctx_create(device, C.byref(h)))
        self.h = h
        self.rank, self.world = rank, world

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = cls(0)
        return cls._default

    def check(self, status: int):
        if status != _ffi.OK:
            msg = lib().zk_status_string(status).decode()
            detail = lib().zk_last_error(self.h).decode()
            raise ZkError(status, msg if not detail.startswith(msg) or status < 12 else detail)

    def stream_ptr(self) -> int:
        return int(lib().zk_ctx_stream(self.h) or 0)

    def synchronize(self):
        self.check(lib().zk_ctx_synchronize(self.h))

    def launch_count(self) -> int:
        return int(lib().zk_ctx_launch_count(self.h))

    def last_round_ms(self) -> List[float]:
        buf = (C.c_float * 600)()
        n = lib().zk_ctx_last_round_ms(self.h, buf, 600)
        return [float(buf[i]) for i in range(n)]

    def last_prove_ms(self):
        buf = (C.c_double * 3)()
        lib().zk_ctx_last_prove_ms(self.h, buf)
        return {"total_ms": buf[0], "absorb_ms": buf[1], "kernel_ms": buf[2]}

    def uses_mailbox(self) -> bool:
        """True when this sharded context all-reduces the round sums inside the reducing launch (peer mailboxes over NVLink)."""
        return bool(lib().zk_ctx_uses_mailbox(self.h))

    def set_gather_threshold(self, n: int):
        self.check(lib().zk_ctx_set_gather_threshold(self.h, n))

    def microbench(self, field: int = BLS12_381_FR) -> dict:
        mb = _ffi.zk_microbench()
        self.check(lib().zk_microbench_run(self.h, field, C.byref(mb)))
        return {n: getattr(mb, n) for n, _ in mb._fields_}

    def close(self):
        if self.h:
            lib().zk_ctx_destroy(self.h)
            self.h = None


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().zk_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def _table_array(tables: Sequence["MultiLinearPolynomial"]):
    arr = (C.c_void_p * max(1, len(tables)))()
    for i, t in enumerate(tables):
        arr[i] = t._h
    return arr


# ---- polynomial::multilinear::evaluation_form::MultiLinearPolynomial ------------------------------------
class MultiLinearPolynomial:
    """Dense evaluation-form MLE resident on the GPU (polynomial/src/multilinear/evaluation_form.rs:7-10)."""

    def __init__(self, handle, ctx: Context):
        self._h = handle
        self.ctx = ctx

    @classmethod
    def new(cls, n_vars: int, evaluations, field: int = BLS12_381_FR, ctx: Optional[Context] = None):
        """new(n_vars, evaluations) :15-27.  `evaluations`: ints, or an (N,4) uint64 Montgomery array."""
        ctx = ctx or Context.default()
        if isinstance(evaluations, np.ndarray):
            mont = np.ascontiguousarray(evaluations, dtype=np.uint64).reshape(-1, 4)
        else:
            mont = to_mont(field, list(evaluations)) if len(evaluations) else np.zeros((0, 4), dtype=np.uint64)
        h = C.c_void_p()
        ctx.check(lib().zk_table_upload(ctx.h, field, mont.ctypes.data if mont.size else None, mont.shape[0], n_vars, C.byref(h)))
        return cls(h, ctx)

    @classmethod
    def new_local(cls, n_vars: int, local_evaluations: np.ndarray, field: int = BLS12_381_FR, ctx: Optional[Context] = None):
        """This rank's local entries uploaded as they are (zk_table_upload_local): (L,4) uint64 Montgomery limbs."""
        ctx = ctx or Context.default()
        mont = np.ascontiguousarray(local_evaluations, dtype=np.uint64).reshape(-1, 4)
        h = C.c_void_p()
        ctx.check(lib().zk_table_upload_local(ctx.h, field, mont.ctypes.data if mont.size else None, mont.shape[0], n_vars, C.byref(h)))
        return cls(h, ctx)

    def ntt_sharded(self, inverse: bool = False):
        """Multi-GPU fft / ifft in place (zk_ntt_sharded): strided shard -> contiguous block (forward), the reverse
        for the inverse.  Collective over the sharded context."""
        self.ctx.check(lib().zk_ntt_sharded(self.ctx.h, self._h, int(inverse)))

    def ntt_virtual_sharded(self, ranks: int, inverse: bool = False):
        """The multi-GPU factorisation with `ranks` virtual ranks on this one GPU (zk_ntt_virtual_sharded)."""
        self.ctx.check(lib().zk_ntt_virtual_sharded(self.ctx.h, self._h, ranks, int(inverse)))

    def ntt(self, inverse: bool = False):
        """fft / ifft of the table in place (zk_ntt), natural order in and out."""
        self.ctx.check(lib().zk_ntt(self.ctx.h, self._h, int(inverse)))

    @classmethod
    def generate(cls, n_vars: int, table_id: int, seed: int = DEFAULT_SEED, field: int = BLS12_381_FR,
                 ctx: Optional[Context] = None):
        """Deterministic synthetic table, generated on the device."""
        ctx = ctx or Context.default()
        h = C.c_void_p()
        ctx.check(lib().zk_table_generate(ctx.h, field, seed, table_id, n_vars, C.byref(h)))
        return cls(h, ctx)

    def regenerate(self, table_id: int, seed: int = DEFAULT_SEED):
        """Refill this table in place with synthetic table `table_id` (reuses the allocation)."""
        self.ctx.check(lib().zk_table_regenerate(self.ctx.h, self._h, seed, table_id))

    def n_vars(self) -> int:  # :30
        return int(lib().zk_table_n_vars(self._h))

    @property
    def field(self) -> int:
        return int(lib().zk_table_field(self._h))

    def local_len(self) -> int:
        return int(lib().zk_table_local_len(self._h))

    def partial_evaluate(self, initial_var: int, assignments: Sequence[int]) -> "MultiLinearPolynomial":  # :40-80
        a = to_mont(self.field, list(assignments)) if len(assignments) else np.zeros((0, 4), dtype=np.uint64)
        h = C.c_void_p()
        self.ctx.check(lib().zk_mle_partial_evaluate(self.ctx.h, self._h, initial_var, a.ctypes.data if a.size else None,
                                                     a.shape[0], C.byref(h)))
        return MultiLinearPolynomial(h, self.ctx)

    def evaluate(self, assignments: Sequence[int]) -> int:  # :83-89
        a = to_mont(self.field, list(assignments)) if len(assignments) else np.zeros((0, 4), dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().zk_mle_evaluate(self.ctx.h, self._h, a.ctypes.data if a.size else None, a.shape[0], out.ctypes.data))
        return from_mont(self.field, out)[0]

    def evaluation_slice_mont(self) -> np.ndarray:
        out = np.empty((self.local_len(), 4), dtype=np.uint64)
        self.ctx.check(lib().zk_table_download(self.ctx.h, self._h, out.ctypes.data))
        return out

    def evaluation_slice(self) -> List[int]:  # :92-94
        return from_mont(self.field, self.evaluation_slice_mont())

    @property
    def evaluations(self) -> List[int]:
        return self.evaluation_slice()

    def to_bytes(self) -> bytes:  # :97-103
        out = np.empty(self.local_len() * 32, dtype=np.uint8)
        self.ctx.check(lib().zk_mle_to_bytes(self.ctx.h, self._h, out.ctypes.data))
        return out.tobytes()

    def clone(self) -> "MultiLinearPolynomial":
        h = C.c_void_p()
        self.ctx.check(lib().zk_table_clone(self.ctx.h, self._h, C.byref(h)))
        return MultiLinearPolynomial(h, self.ctx)

    def __eq__(self, other):  # derive(PartialEq)
        return self.n_vars() == other.n_vars() and (self.evaluation_slice_mont() == other.evaluation_slice_mont()).all()

    def __del__(self):
        try:
            if self._h:
                lib().zk_table_free(self._h)
                self._h = None
        except Exception:
            pass


# ---- polynomial::product_poly::ProductPoly -------------------------------------------------------------
class ProductPoly:
    """P(x) = A(x).B(x).C(x)  (polynomial/src/product_poly.rs:4-10)."""

    def __init__(self, polynomials: Sequence[MultiLinearPolynomial], _checked: bool = False):
        # The reference's ProductPoly OWNS its factors (`ProductPoly::new(vec![f.clone(), f.clone()])` holds two
        # vectors).  Here factors are handles to device tables, and the in-place folds of the prover need one buffer per
        # factor: the same handle listed again is cloned, so ProductPoly.new([f, f]) is f * f like in the reference.
        self.polynomials, seen = [], set()
        for q in polynomials:
            key = q._h.value if hasattr(q._h, "value") else q._h
            self.polynomials.append(q.clone() if key in seen else q)
            seen.add(key)
        self.ctx = self.polynomials[0].ctx if self.polynomials else Context.default()

    @classmethod
    def new(cls, polynomials: Sequence[MultiLinearPolynomial]) -> "ProductPoly":  # :14-32
        polys = list(polynomials)
        _check(lib().zk_product_check(_table_array(polys), len(polys)))
        return cls(polys)

    def _arr(self):
        return _table_array(self.polynomials)

    def n_vars(self) -> int:  # :86-88
        return self.polynomials[0].n_vars()

    @property
    def field(self) -> int:
        return self.polynomials[0].field

    def evaluate(self, assignments: Sequence[int]) -> int:  # :36-44
        a = to_mont(self.field, list(assignments)) if len(assignments) else np.zeros((0, 4), dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().zk_product_evaluate(self.ctx.h, self._arr(), len(self.polynomials),
                                                 a.ctypes.data if a.size else None, a.shape[0], out.ctypes.data))
        return from_mont(self.field, out)[0]

    def partial_evaluate(self, initial_var: int, assignments: Sequence[int]) -> "ProductPoly":  # :48-63
        return ProductPoly([q.partial_evaluate(initial_var, assignments) for q in self.polynomials])

    def prod_reduce(self) -> List[int]:  # :66-74
        h = C.c_void_p()
        self.ctx.check(lib().zk_product_prod_reduce(self.ctx.h, self._arr(), len(self.polynomials), C.byref(h)))
        return MultiLinearPolynomial(h, self.ctx).evaluation_slice()

    def sum_mont(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().zk_product_sum(self.ctx.h, self._arr(), len(self.polynomials), out.ctypes.data))
        return out

    def sum(self) -> int:
        """sum over the hypercube of the product (prod_reduce().iter().sum())."""
        return from_mont(self.field, self.sum_mont())[0]

    def round_poly(self, degree: int) -> List[int]:  # sumcheck/src/prover.rs:48-56
        out = np.zeros((degree + 1, 4), dtype=np.uint64)
        self.ctx.check(lib().zk_product_round_poly(self.ctx.h, self._arr(), len(self.polynomials), degree, out.ctypes.data))
        return from_mont(self.field, out)

    def fold_inplace(self, r: int):  # prover.rs:64
        rm = to_mont(self.field, [r])
        self.ctx.check(lib().zk_product_fold_inplace(self.ctx.h, self._arr(), len(self.polynomials), rm.ctypes.data))

    def fold_then_round_poly(self, r: int, degree: int) -> List[int]:
        rm = to_mont(self.field, [r])
        out = np.zeros((degree + 1, 4), dtype=np.uint64)
        self.ctx.check(lib().zk_product_fold_then_round_poly(self.ctx.h, self._arr(), len(self.polynomials), degree,
                                                             rm.ctypes.data, out.ctypes.data))
        return from_mont(self.field, out)

    def to_bytes(self) -> bytes:  # :77-83
        return b"".join(q.to_bytes() for q in self.polynomials)

    def clone(self) -> "ProductPoly":  # derive(Clone)
        return ProductPoly([q.clone() for q in self.polynomials])

    def __eq__(self, other):
        return len(self.polynomials) == len(other.polynomials) and all(a == b for a, b in zip(self.polynomials, other.polynomials))


# ---- sum of products (beyond the reference's ProductPoly) ------------------------------------
class SumOfProductsPoly:
    """P(x) = sum_t prod_{k in terms[t]} polynomials[k](x) — e.g. the GKR layer polynomial add.Wb + add.Wc + mul.Wb.Wc
    with polynomials = [add, mul, Wb, Wc] and terms = [[0, 2], [0, 3], [1, 2, 3]].  Same surface as ProductPoly
    (the four methods the prover loop of sumcheck/src/prover.rs:33-73 calls, plus evaluate for the verifier's final
    check); a single term listing every table once is the reference's ProductPoly."""

    def __init__(self, polynomials: Sequence[MultiLinearPolynomial], terms: Sequence[Sequence[int]]):
        self.polynomials = list(polynomials)
        self.terms = [list(t) for t in terms]
        self.ctx = self.polynomials[0].ctx if self.polynomials else Context.default()
        self._term_len = np.array([len(t) for t in self.terms], dtype=np.uint8)
        self._term_fac = np.array([k for t in self.terms for k in t] or [0], dtype=np.uint8)

    @classmethod
    def new(cls, polynomials: Sequence[MultiLinearPolynomial], terms: Sequence[Sequence[int]]) -> "SumOfProductsPoly":
        polys = list(polynomials)
        _check(lib().zk_product_check(_table_array(polys), len(polys)))
        if len(terms) == 0 or any(len(t) == 0 for t in terms):
            raise ZkError(3, lib().zk_status_string(3).decode())
        if len(terms) > 8 or any(len(t) > 8 for t in terms):
            raise ZkError(13, lib().zk_status_string(13).decode())
        if any(not (0 <= int(k) < len(polys)) for t in terms for k in t):
            raise ZkError(12, lib().zk_status_string(12).decode())
        return cls(polys, terms)

    def _arr(self):
        return _table_array(self.polynomials)

    def _terms(self):
        return self._term_len.ctypes.data, self._term_fac.ctypes.data, len(self.terms)

    def n_vars(self) -> int:
        return self.polynomials[0].n_vars()

    @property
    def field(self) -> int:
        return self.polynomials[0].field

    def partial_evaluate(self, initial_var: int, assignments: Sequence[int]) -> "SumOfProductsPoly":
        return SumOfProductsPoly([q.partial_evaluate(initial_var, assignments) for q in self.polynomials], self.terms)

    def evaluate(self, assignments: Sequence[int]) -> int:
        a = to_mont(self.field, list(assignments)) if len(assignments) else np.zeros((0, 4), dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().zk_sop_evaluate(self.ctx.h, self._arr(), len(self.polynomials), *self._terms(),
                                             a.ctypes.data if a.size else None, a.shape[0], out.ctypes.data))
        return from_mont(self.field, out)[0]

    def combine(self, table_values: Sequence[int]) -> int:
        """P at a point from the tables' values there (host only)."""
        v = to_mont(self.field, list(table_values))
        out = np.zeros(4, dtype=np.uint64)
        _check(lib().zk_sop_combine(self.field, *self._terms(), v.ctypes.data, v.shape[0], out.ctypes.data))
        return from_mont(self.field, out)[0]

    def sum_mont(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().zk_sop_sum(self.ctx.h, self._arr(), len(self.polynomials), *self._terms(), out.ctypes.data))
        return out

    def sum(self) -> int:
        return from_mont(self.field, self.sum_mont())[0]

    def round_poly(self, degree: int) -> List[int]:
        out = np.zeros((degree + 1, 4), dtype=np.uint64)
        self.ctx.check(lib().zk_sop_round_poly(self.ctx.h, self._arr(), len(self.polynomials), *self._terms(), degree,
                                               out.ctypes.data))
        return from_mont(self.field, out)

    def to_bytes(self) -> bytes:
        return b"".join(q.to_bytes() for q in self.polynomials)

    def clone(self) -> "SumOfProductsPoly":
        return SumOfProductsPoly([q.clone() for q in self.polynomials], self.terms)


# ---- sumcheck ---------------------------------------------------------------------------------------------
class SumcheckProof:
    """sumcheck/src/lib.rs:8-11."""

    def __init__(self, field: int, sum_: int, round_polys: List[List[int]], round_polys_mont: np.ndarray, sum_mont: np.ndarray):
        self.field = field
        self.sum = sum_
        self.round_polys = round_polys
        self._round_polys_mont = round_polys_mont
        self._sum_mont = sum_mont

    def to_bytes(self, challenges: Optional[Sequence[int]] = None, final_evals: Optional[Sequence[int]] = None) -> bytes:
        """Proof dump: BE32(sum) || round polynomials || challenges || final evaluations."""
        rp = np.ascontiguousarray(self._round_polys_mont)
        n = rp.shape[0]
        d1 = rp.shape[1] if rp.ndim == 3 else 1
        ch = to_mont(self.field, list(challenges)) if challenges is not None and len(challenges) else None
        fe = to_mont(self.field, list(final_evals)) if final_evals is not None and len(final_evals) else None
        ln = C.c_size_t(0)
        args = (self.field, self._sum_mont.ctypes.data, rp.ctypes.data if rp.size else None, n, d1 - 1,
                ch.ctypes.data if ch is not None else None, fe.ctypes.data if fe is not None else None,
                fe.shape[0] if fe is not None else 0)
        _check(lib().zk_sumcheck_proof_dump(*args, None, 0, C.byref(ln), None))
        out = np.zeros(ln.value, dtype=np.uint8)
        _check(lib().zk_sumcheck_proof_dump(*args, out.ctypes.data, ln.value, C.byref(ln), None))
        return out.tobytes()

    @classmethod
    def from_values(cls, field: int, sum_: int, round_polys: List[List[int]]):
        d1 = len(round_polys[0]) if round_polys else 1
        flat = [x for rp in round_polys for x in rp]
        rm = to_mont(field, flat).reshape(len(round_polys), d1, 4) if flat else np.zeros((0, d1, 4), dtype=np.uint64)
        return cls(field, sum_ % MODULUS[field], round_polys, rm, to_mont(field, [sum_])[0])


class SubClaim:
    """sumcheck/src/lib.rs:17-20."""

    def __init__(self, sum_: int, challenges: List[int]):
        self.sum = sum_
        self.challenges = challenges


class SumcheckProver:
    """SumcheckProver::<MAX_VAR_DEGREE, F>  (sumcheck/src/prover.rs:9-74)."""

    def __init__(self, max_var_degree: int):
        self.max_var_degree = max_var_degree

    def _run(self, poly: ProductPoly, sum_: int, absorb: bool):
        field, n, m, d1 = poly.field, poly.n_vars(), len(poly.polynomials), self.max_var_degree + 1
        sm = to_mont(field, [sum_])[0]
        rp = np.zeros((n, d1, 4), dtype=np.uint64)
        ch = np.zeros((n, 4), dtype=np.uint64)
        fin = np.zeros((m, 4), dtype=np.uint64)
        if isinstance(poly, SumOfProductsPoly):
            poly.ctx.check(lib().zk_sumcheck_prove_sop(poly.ctx.h, poly._arr(), m, *poly._terms(), self.max_var_degree,
                                                       sm.ctypes.data, int(absorb), rp.ctypes.data, ch.ctypes.data,
                                                       fin.ctypes.data))
        else:
            poly.ctx.check(lib().zk_sumcheck_prove(poly.ctx.h, poly._arr(), m, self.max_var_degree, sm.ctypes.data, int(absorb),
                                                   rp.ctypes.data, ch.ctypes.data, fin.ctypes.data))
        vals = from_mont(field, rp.reshape(-1, 4)) if n else []
        rps = [vals[i * d1 : (i + 1) * d1] for i in range(n)]
        proof = SumcheckProof(field, sum_ % MODULUS[field], rps, rp, sm)
        self.final_evals = from_mont(field, fin)
        return proof, (from_mont(field, ch) if n else [])

    def prove(self, poly: ProductPoly, sum_: int) -> SumcheckProof:  # :15-20  (consumes `poly`)
        return self._run(poly, sum_, True)[0]

    def prove_partial(self, poly: ProductPoly, sum_: int):  # :24-30
        return self._run(poly, sum_, False)


class SumcheckVerifier:
    """sumcheck/src/verifier.rs:9-79."""

    @staticmethod
    def verify(poly: ProductPoly, proof: SumcheckProof) -> bool:  # :15-33
        rp = np.ascontiguousarray(proof._round_polys_mont)
        d1 = rp.shape[1] if rp.ndim == 3 else 1
        if isinstance(poly, SumOfProductsPoly):
            st = lib().zk_sumcheck_verify_sop(poly.ctx.h, poly._arr(), len(poly.polynomials), *poly._terms(), proof._sum_mont.ctypes.data,
                                              rp.ctypes.data if rp.size else None, rp.shape[0], d1 - 1)
        else:
            st = lib().zk_sumcheck_verify(poly.ctx.h, poly._arr(), len(poly.polynomials), proof._sum_mont.ctypes.data,
                                          rp.ctypes.data if rp.size else None, rp.shape[0], d1 - 1)
        if st == 8:  # ZK_VERIFY_FALSE == Ok(false)
            return False
        poly.ctx.check(st)
        return True

    @staticmethod
    def verify_partial(proof: SumcheckProof) -> SubClaim:  # :38-41
        rp = np.ascontiguousarray(proof._round_polys_mont)
        d1 = rp.shape[1] if rp.ndim == 3 else 1
        sub = np.zeros(4, dtype=np.uint64)
        ch = np.zeros((max(1, rp.shape[0]), 4), dtype=np.uint64)
        _check(lib().zk_sumcheck_verify_partial(proof.field, proof._sum_mont.ctypes.data, rp.ctypes.data if rp.size else None,
                                                rp.shape[0], d1 - 1, sub.ctypes.data, ch.ctypes.data))
        return SubClaim(from_mont(proof.field, sub)[0], from_mont(proof.field, ch[: rp.shape[0]]) if rp.shape[0] else [])


# ---- transcript ---------------------------------------------------------------------------------------------
class Transcript:
    """transcript/src/lib.rs:5-35 (host Keccak-256)."""

    def __init__(self):  # new :10
        self._h = C.c_void_p(lib().zk_transcript_new())

    def append(self, new_data: bytes):  # :16
        lib().zk_transcript_append(self._h, bytes(new_data), len(new_data))

    def sample_field_element(self, field: int = BLS12_381_FR) -> int:  # :27
        out = np.zeros(4, dtype=np.uint64)
        _check(lib().zk_transcript_sample_field_element(self._h, field, out.ctypes.data))
        return from_mont(field, out)[0]

    def sample_n_field_elements(self, n: int, field: int = BLS12_381_FR) -> List[int]:  # :32
        return [self.sample_field_element(field) for _ in range(n)]

    def __del__(self):
        try:
            lib().zk_transcript_free(self._h)
        except Exception:
            pass


def keccak256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().zk_keccak256(bytes(data), len(data), C.cast(out, C.c_void_p))
    return out.raw


# ---- fft ---------------------------------------------------------------------------------------------------
def _ntt(values, field: int, inverse: bool, ctx: Optional[Context]):
    ctx = ctx or Context.default()
    as_array = isinstance(values, np.ndarray)
    mont = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4).copy() if as_array else to_mont(field, list(values))
    n = mont.shape[0]
    if n == 0 or n & (n - 1):
        raise ZkError(9, "values must be a power of 2")  # fft/src/lib.rs:29
    ctx.check(lib().zk_ntt_host(ctx.h, field, mont.ctypes.data, n, int(inverse)))
    return mont if as_array else from_mont(field, mont)


def fft(coefficients, field: int = BLS12_381_FR, ctx: Optional[Context] = None):
    """fft/src/lib.rs:4-8: DFT over the 2^k-th roots of unity, natural order in and out."""
    return _ntt(coefficients, field, False, ctx)


def ifft(evaluations, field: int = BLS12_381_FR, ctx: Optional[Context] = None):
    """fft/src/lib.rs:11-19."""
    return _ntt(evaluations, field, True, ctx)


def bind_host_to_gpu(device: int = 0):
    """Pin the calling process to the CPU cores NVML reports as local to `device` (its PCIe root / NUMA node), so that
    pinned host buffers allocated afterwards live in memory next to the GPU's PCIe link.  Host-to-device copies of the
    tables are the whole cost of the host-buffer entry point (`zk_sumcheck_prove_host`), and a buffer on the far socket
    costs 20-25 % of the PCIe rate.  Returns the list of cores, or None when NVML / the topology is unavailable (then
    nothing is changed).  Plain deployment hygiene (`numactl --cpunodebind` by hand does the same)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (dom, bus, dev)).encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1}
        use = local & os.sched_getaffinity(0)
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return sorted(use)
    except Exception:
        return None
