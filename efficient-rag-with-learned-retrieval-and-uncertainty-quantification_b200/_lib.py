"""ctypes binding of libragb200.so (the C ABI declared in include/ragb200.h).

There is no fallback: if the shared library is missing this module raises at import
time, and every entry point refuses devices that are not sm_100.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
# RAGB_LIB_NAME: diagnostics only (an instrumented build of the same sources, e.g. libragb200_prof.so)
LIB_PATH = PKG_DIR / os.environ.get("RAGB_LIB_NAME", "libragb200.so")

RAGB_OK, RAGB_EINVAL, RAGB_EARCH, RAGB_ECUDA, RAGB_ELIMIT, RAGB_ENOSPC = 0, -1, -2, -3, -4, -5
MAX_TOPK = 256
MAX_QUERY_TERMS = 64
GEMV_MAX_BATCH = 8
MMA_MAX_TOPK = 100
FUSE_MAX_POOL = 256
ROUTER_MAX_HIDDEN = 128


class RagbError(RuntimeError):
    """A libragb200 call failed; ``code`` is the RAGB_E* status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libragb200 error {code}: {message}")
        self.code = code


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). rag_uq_b200 has no CPU or generic-GPU fallback."
        )
    return C.CDLL(str(LIB_PATH))


lib = _load()

_p, _i32, _i64, _u64, _f64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/ragb200.h one to one
SIGNATURES = {
    "ragb_abi_version": (C.c_int, []),
    "ragb_last_error": (C.c_char_p, []),
    "ragb_device_check": (C.c_int, [C.c_int]),
    "ragb_launch_count": (_i64, []),
    "ragb_bm25_idf_scratch_bytes": (_sz, [_i64]),
    "ragb_bm25_build_idf": (C.c_int, [_p, _i64, _i64, _f64, _p, _p, _sz, _p]),
    "ragb_bm25_build_norm": (C.c_int, [_p, _i64, _f64, _f64, _f64, _p, _p]),
    "ragb_bm25_term_max_tf": (C.c_int, [_p, _p, _p, _i32, _p, _p]),
    "ragb_bm25_build_dense_table": (C.c_int, [_p, _p, _p, _p, _i32, _i64, _p, _i64, _p]),
    "ragb_bm25_build_impact_bounds": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p, _p]),
    "ragb_bm25_build_posting_impacts": (C.c_int, [_p, _p, _p, _i64, _p, _p]),
    "ragb_bm25_topk_workspace_bytes": (_sz, [_i32, _i64, _i32]),
    "ragb_bm25_score_topk": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f64, _p, _i64, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32,
                                       _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "ragb_bm25_stripe_count": (_i32, [_i32, _i64]),
    "ragb_bm25_score_part": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f64, _p, _i64, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32,
                                       _i64, _i64, _i32, _p, _i32, _i32, _i64, _p, _sz, _p]),
    "ragb_bm25_score_finish": (C.c_int, [_i32, _i64, _i32, _p, _p, _p, _sz, _p]),
    "ragb_bm25_seed": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f64, _p, _i64, _p, _i32, _p, _p, _i32, _i32, _i64, _i32, _p, _p]),
    "ragb_bm25_scores": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f64, _p, _i64, _p, _i32, _p, _p, _i32, _i32, _i64,
                                   _p, _i64, _i32, _p]),
    "ragb_dense_gemv_workspace_bytes": (_sz, [_i32, _i32]),
    "ragb_dense_gemv_topk": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i64, _p, _p, _p, _sz, _p]),
    "ragb_dense_mma_workspace_bytes": (_sz, [_i32, _i32]),
    "ragb_dense_mma_topk": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i64, _i32, _p, _p, _p, _sz, _p]),
    "ragb_dense_mma_topk_min": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "ragb_dense_mma_sample": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i64, _i32, _p, _p, _sz, _p]),
    "ragb_dense_mma_seeded": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "ragb_dense_mma_fused_topk": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i64, _p, _i64, _p, _p, _p, _p, _p, _i32,
                                            _p, _i32, _i32, C.c_float, C.c_float, _p, _p, _p, _p, _sz, _p]),
    "ragb_bm25_score_docs": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f64, _p, _i64, _p, _i32, _p, _p, _i32, _i32, _i64, _i64,
                                       _p, _i32, _p, _p]),
    "ragb_dense_score_docs": (C.c_int, [_p, _i64, _i32, _p, _i32, _i64, _p, _i32, _p, _p]),
    "ragb_dense_scores": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _p]),
    "ragb_topk_rows_workspace_bytes": (_sz, [_i32, _i64, _i32]),
    "ragb_topk_rows": (C.c_int, [_p, _i32, _i64, _i32, _p, _p, _p, _sz, _p]),
    "ragb_topk_merge": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "ragb_topk_merge_strided": (C.c_int, [_p, _p, _i32, _i32, _i32, _i64, _i64, _i32, _p, _p, _p]),
    "ragb_hybrid_fuse_topk": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "ragb_retrieval_uncertainty": (C.c_int, [_p, _p, _i32, _i32, _f64, _p, _p]),
    "ragb_router_scratch_bytes": (_sz, [_i32, _i32]),
    "ragb_router_forward": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p]),
    "ragb_router_mc_dropout": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _p, _i32, _i32, _i32, _f64, _u64, _u64,
                                         _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.ragb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a status code to the Python exception a caller of the reference would expect."""
    if rc == RAGB_OK:
        return
    msg = last_error()
    if rc in (RAGB_EINVAL, RAGB_ELIMIT):
        raise ValueError(f"libragb200: {msg}")
    raise RagbError(rc, msg)


def launch_count() -> int:
    return int(lib.ragb_launch_count())
