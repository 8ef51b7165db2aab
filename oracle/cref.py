"""ctypes loader for oracle/libzkoracle.so (the C restatement in oracle/cpu_ref.c).

TEST INFRASTRUCTURE ONLY — see the header of cpu_ref.c.  Elements cross this boundary as numpy
uint64 arrays of shape (count, 4): little-endian limbs, Montgomery form, fully reduced.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzkoracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libzkoracle.so"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.zko_prove.restype = C.c_int
        _lib.zko_prove_fast.restype = C.c_int
        _lib.zko_prove_fast_mt.restype = C.c_int
        _lib.zko_prove_generated_mt.restype = C.c_int
        _lib.zko_prove_sop.restype = C.c_int
        _lib.zko_sop_sum.restype = C.c_int
        _lib.zko_verify_internal.restype = C.c_int
        _lib.zko_fft.restype = C.c_int
        _lib.zko_fft_fast.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _ptr_array(tables):
    arr = (u64p * len(tables))()
    for i, t in enumerate(tables):
        arr[i] = _p(t)
    return arr


def keccak256(data: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    lib().zko_keccak256(buf, C.c_size_t(len(data)), out)
    return bytes(out)


def ints_to_limbs(vals) -> np.ndarray:
    """Python ints (< 2^256) -> (count,4) uint64 little-endian limbs (no Montgomery conversion)."""
    raw = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(raw, dtype="<u8").reshape(-1, 4).copy()


def limbs_to_ints(a: np.ndarray):
    raw = np.ascontiguousarray(a, dtype="<u8").tobytes()
    return [int.from_bytes(raw[32 * i : 32 * i + 32], "little") for i in range(len(raw) // 32)]


def to_mont(field: int, canon: np.ndarray) -> np.ndarray:
    out = np.empty_like(canon)
    lib().zko_to_mont(field, _p(canon), _p(out), C.c_size_t(canon.shape[0]))
    return out


def from_mont(field: int, mont: np.ndarray) -> np.ndarray:
    mont = np.ascontiguousarray(mont)
    out = np.empty_like(mont)
    lib().zko_from_mont(field, _p(mont), _p(out), C.c_size_t(mont.shape[0]))
    return out


def ints_to_mont(field: int, vals) -> np.ndarray:
    return to_mont(field, ints_to_limbs(vals))


def mont_to_ints(field: int, mont: np.ndarray):
    return limbs_to_ints(from_mont(field, np.ascontiguousarray(mont).reshape(-1, 4)))


def gen_table(field: int, seed: int, table_id: int, n_vars: int, first: int = 0, stride: int = 1, count=None) -> np.ndarray:
    if count is None:
        count = 1 << n_vars
    out = np.empty((count, 4), dtype=np.uint64)
    lib().zko_gen_table(field, C.c_uint64(seed), C.c_uint64(table_id), n_vars, C.c_uint64(first), C.c_uint64(stride),
                        C.c_uint64(count), _p(out))
    return out


def partial_evaluate(field: int, evals: np.ndarray, n_vars: int, initial_var: int, assign: np.ndarray) -> np.ndarray:
    assign = np.ascontiguousarray(assign).reshape(-1, 4)
    out = np.empty((1 << (n_vars - assign.shape[0]), 4), dtype=np.uint64)
    lib().zko_partial_evaluate(field, _p(evals), n_vars, initial_var, _p(assign), assign.shape[0], _p(out))
    return out


def product_sum(field: int, tables, n_vars: int) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64)
    lib().zko_product_sum(field, _ptr_array(tables), len(tables), n_vars, _p(out))
    return out


def prove(field: int, tables, n_vars: int, degree: int, sum_mont: np.ndarray, absorb: bool, fast: bool = False, threads: int = 0):
    """Returns (round_polys (n_vars, degree+1, 4), challenges (n_vars, 4), finals (m, 4)), Montgomery limbs."""
    m = len(tables)
    rp = np.zeros((n_vars, degree + 1, 4), dtype=np.uint64)
    ch = np.zeros((n_vars, 4), dtype=np.uint64)
    fin = np.zeros((m, 4), dtype=np.uint64)
    sum_mont = np.ascontiguousarray(sum_mont, dtype=np.uint64)
    if fast:
        work = [t.copy() for t in tables]
        if threads > 0:  # the streamlined prover on `threads` cores (pthreads), bit-identical
            assert not absorb
            rc = lib().zko_prove_fast_mt(field, _ptr_array(work), m, n_vars, degree, _p(sum_mont), _p(rp), _p(ch), _p(fin), int(threads))
        else:
            rc = lib().zko_prove_fast(field, _ptr_array(work), m, n_vars, degree, _p(sum_mont), int(bool(absorb)), _p(rp), _p(ch), _p(fin))
    else:
        rc = lib().zko_prove(field, _ptr_array(tables), m, n_vars, degree, _p(sum_mont), int(bool(absorb)), _p(rp), _p(ch),
                             _p(fin))
    assert rc == 0
    return rp, ch, fin


def prove_generated(field: int, seed: int, m: int, n_vars: int, degree: int, threads: int = 1):
    """prove_partial of the seeded generator tables 0..m-1 (claim = their true sum) without materialising them at full
    size (zko_prove_generated_mt).  Returns (claim (4,), round_polys, challenges, finals)."""
    claim = np.zeros(4, dtype=np.uint64)
    rp = np.zeros((n_vars, degree + 1, 4), dtype=np.uint64)
    ch = np.zeros((n_vars, 4), dtype=np.uint64)
    fin = np.zeros((m, 4), dtype=np.uint64)
    rc = lib().zko_prove_generated_mt(field, C.c_uint64(seed), m, n_vars, degree, _p(claim), _p(rp), _p(ch), _p(fin), int(threads))
    assert rc == 0, rc
    return claim, rp, ch, fin


def _terms_flat(terms):
    """terms: list of lists of table indices -> (term_len u8[n_terms], term_fac u8[n_terms*8], packed u8[sum len])."""
    tl = np.array([len(t) for t in terms], dtype=np.uint8)
    tf = np.zeros((len(terms), 8), dtype=np.uint8)
    for i, t in enumerate(terms):
        tf[i, : len(t)] = t
    return tl, np.ascontiguousarray(tf.reshape(-1))


def sop_sum(field: int, tables, terms, n_vars: int) -> np.ndarray:
    tl, tf = _terms_flat(terms)
    out = np.empty(4, dtype=np.uint64)
    lib().zko_sop_sum(field, _ptr_array(tables), len(tables), n_vars, tl.ctypes.data_as(u8p), tf.ctypes.data_as(u8p),
                      len(terms), _p(out))
    return out


def prove_sop(field: int, tables, terms, n_vars: int, degree: int, sum_mont: np.ndarray, absorb: bool = False):
    """Sum-of-products prover.  Returns (round_polys (n,degree+1,4), challenges (n,4), finals (n_tables,4))."""
    tl, tf = _terms_flat(terms)
    rp = np.zeros((n_vars, degree + 1, 4), dtype=np.uint64)
    ch = np.zeros((n_vars, 4), dtype=np.uint64)
    fin = np.zeros((len(tables), 4), dtype=np.uint64)
    sum_mont = np.ascontiguousarray(sum_mont, dtype=np.uint64)
    rc = lib().zko_prove_sop(field, _ptr_array(tables), len(tables), n_vars, tl.ctypes.data_as(u8p), tf.ctypes.data_as(u8p),
                             len(terms), degree, _p(sum_mont), int(bool(absorb)), _p(rp), _p(ch), _p(fin))
    assert rc == 0
    return rp, ch, fin


def verify_internal(field: int, sum_mont, round_polys: np.ndarray, initial_poly_bytes: bytes | None = None):
    """Returns (rc, subclaim_sum (4,), challenges (n,4)); rc 0 = Ok, 3 = round check failed."""
    n_rounds, d1 = round_polys.shape[0], round_polys.shape[1]
    sub = np.zeros(4, dtype=np.uint64)
    ch = np.zeros((n_rounds, 4), dtype=np.uint64)
    rp = np.ascontiguousarray(round_polys, dtype=np.uint64)
    sum_mont = np.ascontiguousarray(sum_mont, dtype=np.uint64)
    if initial_poly_bytes is None:
        buf, n = None, 0
    else:
        buf = (C.c_uint8 * len(initial_poly_bytes)).from_buffer_copy(initial_poly_bytes)
        n = len(initial_poly_bytes)
    rc = lib().zko_verify_internal(field, buf, C.c_size_t(n), _p(sum_mont), _p(rp), n_rounds, d1 - 1, _p(sub), _p(ch))
    return rc, sub, ch


def to_bytes(field: int, evals: np.ndarray) -> bytes:
    evals = np.ascontiguousarray(evals).reshape(-1, 4)
    out = (C.c_uint8 * (32 * evals.shape[0]))()
    lib().zko_to_bytes(field, _p(evals), C.c_size_t(evals.shape[0]), out)
    return bytes(out)


def fft(field: int, data: np.ndarray, log_n: int, inverse: bool = False, fast: bool = False) -> np.ndarray:
    data = np.ascontiguousarray(data, dtype=np.uint64)
    out = np.empty_like(data)
    fn = lib().zko_fft_fast if fast else lib().zko_fft
    rc = fn(field, _p(data), _p(out), log_n, int(bool(inverse)))
    if rc != 0:
        raise ValueError("called `Option::unwrap()` on a `None` value")
    return out
