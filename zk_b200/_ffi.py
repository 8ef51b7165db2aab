"""ctypes binding of include/zk_b200.h (libzk_b200.so).  The library is the product: if it is missing or
no CUDA device is present, calls fail loudly — there is no Python/CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZK_B200_LIB: tuning builds of the same library (scripts/gpu_ab.sh); default = the in-tree product build
SO_PATH = os.environ.get("ZK_B200_LIB") or os.path.join(_HERE, "libzk_b200.so")

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)

OK = 0
STATUS_NAMES = {
    0: "ZK_OK", 1: "ZK_ERR_EVAL_LEN", 2: "ZK_ERR_EVALUATE_ARITY", 3: "ZK_ERR_EMPTY_PRODUCT", 4: "ZK_ERR_NVARS_MISMATCH",
    5: "ZK_ERR_PROOF_ROUNDS", 6: "ZK_ERR_INITIAL_EVAL", 7: "ZK_ERR_ROUND_CHECK", 8: "ZK_VERIFY_FALSE", 9: "ZK_ERR_NOT_POW2",
    10: "ZK_ERR_NO_ROOT", 11: "ZK_ERR_VAR_RANGE", 12: "ZK_ERR_INVALID_ARG", 13: "ZK_ERR_UNSUPPORTED", 14: "ZK_ERR_CUDA",
    15: "ZK_ERR_NCCL", 16: "ZK_ERR_OOM",
}


class zk_microbench(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("imad_wide_per_s", "imad_lo_per_s", "iadd3_per_s", "mixed_per_s", "fe_mul_per_s",
                                          "copy_gbs", "read_gbs", "sm_clock_mhz", "dfma_per_s", "fe_mul_fixed_per_s")]


# name -> (restype, argtypes): every symbol include/zk_b200.h declares
SIGNATURES = {
    "zk_status_string": (C.c_char_p, [C.c_int]),
    "zk_last_error": (C.c_char_p, [vp]),
    "zk_ctx_create": (C.c_int, [C.c_int, vpp]),
    "zk_nccl_unique_id": (C.c_int, [vp]),
    "zk_ctx_create_sharded": (C.c_int, [C.c_int, C.c_int, C.c_int, vp, vpp]),
    "zk_ctx_destroy": (None, [vp]),
    "zk_ctx_rank": (C.c_int, [vp]),
    "zk_ctx_world": (C.c_int, [vp]),
    "zk_ctx_uses_mailbox": (C.c_int, [vp]),
    "zk_ctx_set_gather_threshold": (C.c_int, [vp, C.c_uint64]),
    "zk_ctx_launch_count": (C.c_uint64, [vp]),
    "zk_ctx_last_round_ms": (C.c_uint, [vp, C.POINTER(C.c_float), C.c_uint]),
    "zk_ctx_last_prove_ms": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "zk_ctx_synchronize": (C.c_int, [vp]),
    "zk_ctx_stream": (vp, [vp]),
    "zk_host_alloc": (C.c_int, [C.c_size_t, vpp]),
    "zk_host_free": (C.c_int, [vp]),
    "zk_table_upload": (C.c_int, [vp, C.c_int, vp, C.c_uint64, C.c_uint, vpp]),
    "zk_table_generate": (C.c_int, [vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint, vpp]),
    "zk_table_regenerate": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64]),
    "zk_table_upload_local": (C.c_int, [vp, C.c_int, vp, C.c_uint64, C.c_uint, vpp]),
    "zk_table_clone": (C.c_int, [vp, vp, vpp]),
    "zk_table_free": (None, [vp]),
    "zk_table_n_vars": (C.c_uint, [vp]),
    "zk_table_local_len": (C.c_uint64, [vp]),
    "zk_table_field": (C.c_int, [vp]),
    "zk_table_download": (C.c_int, [vp, vp, vp]),
    "zk_mle_partial_evaluate": (C.c_int, [vp, vp, C.c_uint, vp, C.c_uint, vpp]),
    "zk_mle_evaluate": (C.c_int, [vp, vp, vp, C.c_uint, vp]),
    "zk_mle_to_bytes": (C.c_int, [vp, vp, vp]),
    "zk_product_check": (C.c_int, [vpp, C.c_uint]),
    "zk_product_evaluate": (C.c_int, [vp, vpp, C.c_uint, vp, C.c_uint, vp]),
    "zk_product_prod_reduce": (C.c_int, [vp, vpp, C.c_uint, vpp]),
    "zk_product_sum": (C.c_int, [vp, vpp, C.c_uint, vp]),
    "zk_product_round_poly": (C.c_int, [vp, vpp, C.c_uint, C.c_uint, vp]),
    "zk_product_fold_inplace": (C.c_int, [vp, vpp, C.c_uint, vp]),
    "zk_product_fold_then_round_poly": (C.c_int, [vp, vpp, C.c_uint, C.c_uint, vp, vp]),
    "zk_sumcheck_prove": (C.c_int, [vp, vpp, C.c_uint, C.c_uint, vp, C.c_int, vp, vp, vp]),
    "zk_sumcheck_prove_host": (C.c_int, [vp, C.c_int, vpp, C.c_uint, C.c_uint, C.c_uint, vp, C.c_int, vp, vp, vp, vp]),
    "zk_sumcheck_verify": (C.c_int, [vp, vpp, C.c_uint, vp, vp, C.c_uint, C.c_uint]),
    "zk_sumcheck_verify_partial": (C.c_int, [C.c_int, vp, vp, C.c_uint, C.c_uint, vp, vp]),
    "zk_sumcheck_proof_dump": (C.c_int, [C.c_int, vp, vp, C.c_uint, C.c_uint, vp, vp, C.c_uint, vp, C.c_size_t, C.POINTER(C.c_size_t), vp]),
    "zk_sop_round_poly": (C.c_int, [vp, vpp, C.c_uint, vp, vp, C.c_uint, C.c_uint, vp]),
    "zk_sop_sum": (C.c_int, [vp, vpp, C.c_uint, vp, vp, C.c_uint, vp]),
    "zk_sop_evaluate": (C.c_int, [vp, vpp, C.c_uint, vp, vp, C.c_uint, vp, C.c_uint, vp]),
    "zk_sop_combine": (C.c_int, [C.c_int, vp, vp, C.c_uint, vp, C.c_uint, vp]),
    "zk_sumcheck_prove_sop": (C.c_int, [vp, vpp, C.c_uint, vp, vp, C.c_uint, C.c_uint, vp, C.c_int, vp, vp, vp]),
    "zk_sumcheck_verify_sop": (C.c_int, [vp, vpp, C.c_uint, vp, vp, C.c_uint, vp, vp, C.c_uint, C.c_uint]),
    "zk_transcript_new": (vp, []),
    "zk_transcript_free": (None, [vp]),
    "zk_transcript_append": (None, [vp, C.c_char_p, C.c_size_t]),
    "zk_transcript_sample_field_element": (C.c_int, [vp, C.c_int, vp]),
    "zk_transcript_sample_n_field_elements": (C.c_int, [vp, C.c_int, C.c_uint, vp]),
    "zk_keccak256": (None, [vp, C.c_size_t, vp]),
    "zk_ntt": (C.c_int, [vp, vp, C.c_int]),
    "zk_ntt_host": (C.c_int, [vp, C.c_int, vp, C.c_uint64, C.c_int]),
    "zk_ntt_sharded": (C.c_int, [vp, vp, C.c_int]),
    "zk_ntt_virtual_sharded": (C.c_int, [vp, vp, C.c_uint, C.c_int]),
    "zk_field_from_canonical": (C.c_int, [C.c_int, vp, vp, C.c_size_t]),
    "zk_field_to_canonical": (C.c_int, [C.c_int, vp, vp, C.c_size_t]),
    "zk_field_from_u64": (C.c_int, [C.c_int, C.c_uint64, vp]),
    "zk_field_to_bytes_be": (C.c_int, [C.c_int, vp, C.c_size_t, vp]),
    "zk_field_from_be_bytes_mod_order": (C.c_int, [C.c_int, C.c_char_p, vp]),
    "zk_field_mul": (C.c_int, [C.c_int, vp, vp, vp]),
    "zk_field_add": (C.c_int, [C.c_int, vp, vp, vp]),
    "zk_field_sub": (C.c_int, [C.c_int, vp, vp, vp]),
    "zk_field_inverse": (C.c_int, [C.c_int, vp, vp]),
    "zk_field_root_of_unity": (C.c_int, [C.c_int, C.c_uint64, vp]),
    "zk_round_poly_evaluate": (C.c_int, [C.c_int, vp, C.c_uint, vp, vp]),
    "zk_microbench_run": (C.c_int, [vp, C.c_int, C.POINTER(zk_microbench)]),
}

_lib = None


def lib():
    """Load libzk_b200.so (raises if it has not been built: run `make` or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise OSError(f"{SO_PATH} not found: build it with `make` (there is no fallback implementation)")
        l = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib
