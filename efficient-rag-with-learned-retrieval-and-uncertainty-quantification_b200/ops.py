"""PyTorch custom ops (torch.library) over the C ABI of libragb200.so.

Every op is registered for the CUDA device only (``device_types="cuda"``): calling one with
CPU tensors raises NotImplementedError from the dispatcher - there is no CPU implementation
on purpose.  torch is used for memory, streams and dispatch; all arithmetic happens in the
hand-written sm_100a kernels.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, lib

NS = "rag_uq_b200"


def _ptr(t: Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: Tensor, dtype: torch.dtype, name: str) -> Tensor:
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise NotImplementedError(f"{name}: rag_uq_b200 ops run on CUDA (sm_100) tensors only")
    return t if t.is_contiguous() else t.contiguous()


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------
# BM25
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::bm25_build_idf", mutates_args=(), device_types="cuda")
def bm25_build_idf(df: Tensor, corpus_size: int, epsilon: float) -> Tensor:
    df = _need(df, torch.int32, "df")
    out = torch.empty(df.shape[0], dtype=torch.float32, device=df.device)
    nbytes = lib.ragb_bm25_idf_scratch_bytes(df.shape[0])
    scratch = _workspace(nbytes, df.device)
    with torch.cuda.device(df.device):
        check(lib.ragb_bm25_build_idf(_ptr(df), df.shape[0], corpus_size, epsilon, _ptr(out), _ptr(scratch),
                                      scratch.numel(), _stream()))
    return out


@bm25_build_idf.register_fake
def _(df, corpus_size, epsilon):
    return df.new_empty(df.shape, dtype=torch.float32)


@torch.library.custom_op(f"{NS}::bm25_build_norm", mutates_args=(), device_types="cuda")
def bm25_build_norm(doc_len: Tensor, avgdl: float, k1: float, b: float) -> Tensor:
    doc_len = _need(doc_len, torch.int32, "doc_len")
    out = torch.empty(doc_len.shape[0], dtype=torch.float32, device=doc_len.device)
    with torch.cuda.device(doc_len.device):
        check(lib.ragb_bm25_build_norm(_ptr(doc_len), doc_len.shape[0], avgdl, k1, b, _ptr(out), _stream()))
    return out


@bm25_build_norm.register_fake
def _(doc_len, avgdl, k1, b):
    return doc_len.new_empty(doc_len.shape, dtype=torch.float32)


@torch.library.custom_op(f"{NS}::bm25_term_max_tf", mutates_args=(), device_types="cuda")
def bm25_term_max_tf(term_off: Tensor, post_tf: Tensor, terms: Tensor) -> Tensor:
    """Largest term frequency in the posting list of each of ``terms`` (int32 [n])."""
    term_off, post_tf, terms = _need(term_off, torch.int64, "term_off"), _need(post_tf, torch.int16, "post_tf"), \
        _need(terms, torch.int32, "terms")
    out = torch.empty(terms.shape[0], dtype=torch.int32, device=terms.device)
    with torch.cuda.device(terms.device):
        check(lib.ragb_bm25_term_max_tf(_ptr(term_off), _ptr(post_tf), _ptr(terms), terms.shape[0], _ptr(out), _stream()))
    return out


@bm25_term_max_tf.register_fake
def _(term_off, post_tf, terms):
    return terms.new_empty(terms.shape)


@torch.library.custom_op(f"{NS}::bm25_build_dense_table", mutates_args=(), device_types="cuda")
def bm25_build_dense_table(term_off: Tensor, post_doc: Tensor, post_tf: Tensor, terms: Tensor, n_docs: int,
                           stride: int) -> Tensor:
    """uint8 [n_terms, stride] tf rows of ``terms`` (0 = absent)."""
    term_off, post_doc, post_tf, terms = _need(term_off, torch.int64, "term_off"), _need(post_doc, torch.int32, "post_doc"), \
        _need(post_tf, torch.int16, "post_tf"), _need(terms, torch.int32, "terms")
    table = torch.empty((terms.shape[0], stride), dtype=torch.uint8, device=terms.device)
    with torch.cuda.device(terms.device):
        check(lib.ragb_bm25_build_dense_table(_ptr(term_off), _ptr(post_doc), _ptr(post_tf), _ptr(terms), terms.shape[0],
                                              n_docs, _ptr(table), stride, _stream()))
    return table


@bm25_build_dense_table.register_fake
def _(term_off, post_doc, post_tf, terms, n_docs, stride):
    return terms.new_empty((terms.shape[0], stride), dtype=torch.uint8)


@torch.library.custom_op(f"{NS}::bm25_build_impact_bounds", mutates_args=(), device_types="cuda")
def bm25_build_impact_bounds(dense_tf: Tensor, norm: Tensor) -> Tuple[Tensor, Tensor]:
    """-> (fp16 [n_dense, stride] upper bounds of tf / (tf + norm), fp32 [n_dense] row maxima)."""
    dense_tf, norm = _need(dense_tf, torch.uint8, "dense_tf"), _need(norm, torch.float32, "norm")
    rows, stride = dense_tf.shape
    imp = torch.empty((rows, stride), dtype=torch.float16, device=dense_tf.device)
    maximp = torch.empty(rows, dtype=torch.float32, device=dense_tf.device)
    with torch.cuda.device(dense_tf.device):
        check(lib.ragb_bm25_build_impact_bounds(_ptr(dense_tf), stride, rows, _ptr(norm), norm.shape[0], _ptr(imp),
                                                _ptr(maximp), _stream()))
    return imp, maximp


@bm25_build_impact_bounds.register_fake
def _(dense_tf, norm):
    return dense_tf.new_empty(dense_tf.shape, dtype=torch.float16), norm.new_empty((dense_tf.shape[0],))


def _dense_table(dense_tf: Tensor, dense_terms: Tensor, n_docs: int):
    """(ptr, stride, terms ptr, rows) of the optional dense tf table; an empty tensor disables it."""
    if dense_tf.numel() == 0 or dense_terms.numel() == 0:
        return None, 0, None, 0
    dense_tf = _need(dense_tf, torch.uint8, "dense_tf")
    dense_terms = _need(dense_terms, torch.int32, "dense_terms")
    if dense_tf.dim() != 2 or dense_tf.shape[0] != dense_terms.shape[0] or dense_tf.shape[1] < n_docs:
        raise ValueError("dense_tf must be [n_dense, stride >= n_docs] with one row per entry of dense_terms")
    return dense_tf.data_ptr(), dense_tf.shape[1], dense_terms.data_ptr(), dense_terms.shape[0]


def _impact_cap(dense_cap: Tensor, hi_off: Tensor, hi_doc: Tensor, n_dense: int):
    """(cap ptr, hi_off ptr, hi_doc ptr) of the optional impact cap; empty tensors disable it."""
    if not n_dense or dense_cap.numel() == 0 or hi_off.numel() == 0:
        return None, None, None
    dense_cap, hi_off, hi_doc = _need(dense_cap, torch.float32, "dense_cap"), _need(hi_off, torch.int32, "hi_off"), \
        _need(hi_doc, torch.int32, "hi_doc")
    if dense_cap.shape[0] != n_dense or hi_off.shape[0] != n_dense + 1:
        raise ValueError("dense_cap must hold one cap per table row and hi_off n_dense + 1 offsets")
    return dense_cap.data_ptr(), hi_off.data_ptr(), (hi_doc.data_ptr() if hi_doc.numel() else dense_cap.data_ptr())


def _posting_impacts(post_imp: Tensor, post_doc: Tensor):
    if post_imp is None or post_imp.numel() == 0:
        return None
    post_imp = _need(post_imp, torch.float32, "post_imp")
    if post_imp.shape[0] != post_doc.shape[0]:
        raise ValueError("post_imp must hold one impact per posting")
    return post_imp.data_ptr()


def bm25_build_posting_impacts(post_doc: Tensor, post_tf: Tensor, norm: Tensor) -> Tensor:
    """post_imp[i] = tf / (tf + norm[doc]) of every posting, with the search kernel's own expression (ragb200.h)."""
    post_doc = _need(post_doc, torch.int32, "post_doc")
    post_tf = _need(post_tf, torch.int16, "post_tf")
    norm = _need(norm, torch.float32, "norm")
    out = torch.empty(post_doc.shape[0], dtype=torch.float32, device=post_doc.device)
    if post_doc.shape[0]:
        with torch.cuda.device(post_doc.device):
            check(lib.ragb_bm25_build_posting_impacts(_ptr(post_doc), _ptr(post_tf), _ptr(norm), post_doc.shape[0], _ptr(out),
                                                      _stream()))
    return out


def _bm25_args(term_off, post_doc, post_tf, norm, idf, q_terms, q_off):
    term_off = _need(term_off, torch.int64, "term_off")
    post_doc = _need(post_doc, torch.int32, "post_doc")
    post_tf = _need(post_tf, torch.int16, "post_tf")  # bit pattern of uint16
    norm = _need(norm, torch.float32, "norm")
    idf = _need(idf, torch.float32, "idf")
    q_terms = _need(q_terms, torch.int32, "q_terms")
    q_off = _need(q_off, torch.int32, "q_off")
    if term_off.shape[0] != idf.shape[0] + 1:
        raise ValueError("term_off must have vocab + 1 entries")
    return term_off, post_doc, post_tf, norm, idf, q_terms, q_off


@torch.library.custom_op(f"{NS}::bm25_score_topk", mutates_args=(), device_types="cuda")
def bm25_score_topk(term_off: Tensor, post_doc: Tensor, post_tf: Tensor, norm: Tensor, idf: Tensor, k1: float,
                    dense_tf: Tensor, dense_terms: Tensor, dense_imp: Tensor, dense_maximp: Tensor, q_terms: Tensor,
                    q_off: Tensor, max_query_terms: int, id_base: int, k: int, seed: Tensor, dense_cap: Tensor,
                    hi_off: Tensor, hi_doc: Tensor, post_imp: Tensor) -> Tuple[Tensor, Tensor]:
    """dense_imp (float16 [n_dense, stride]) / dense_maximp (float32 [n_dense]): optional impact bounds of the table
    terms (ragb200.h); pass empty tensors to run without them - the results are the same.
    seed (float32 [B] or empty): proven lower bounds of every query's k-th best score (``bm25_seed``, possibly raised
    to the maximum over all shards); empty = the kernel seeds itself.
    dense_cap (float32 [n_dense]) / hi_off (int32 [n_dense + 1]) / hi_doc (int32): optional impact cap of the table rows
    and the marker lists of the documents above it (ragb200.h); empty tensors = none.  Pruning only.
    post_imp (float32 [nnz] or empty): baked impacts of the postings (``bm25_build_posting_impacts``); speed only."""
    term_off, post_doc, post_tf, norm, idf, q_terms, q_off = _bm25_args(term_off, post_doc, post_tf, norm, idf,
                                                                        q_terms, q_off)
    n_q, n_docs, dev = q_off.shape[0] - 1, norm.shape[0], norm.device
    score = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    ws = _workspace(lib.ragb_bm25_topk_workspace_bytes(n_q, n_docs, k), dev)
    dt, stride, dterms, n_dense = _dense_table(dense_tf, dense_terms, n_docs)
    imp, maximp = None, None
    if n_dense and dense_imp.numel() and dense_maximp.numel():
        dense_imp = _need(dense_imp, torch.float16, "dense_imp")
        dense_maximp = _need(dense_maximp, torch.float32, "dense_maximp")
        if tuple(dense_imp.shape) != tuple(dense_tf.shape) or dense_maximp.shape[0] != n_dense:
            raise ValueError("dense_imp must have the shape of dense_tf and dense_maximp one entry per row")
        imp, maximp = dense_imp.data_ptr(), dense_maximp.data_ptr()
    seed_ptr = None
    if seed.numel():
        seed = _need(seed, torch.float32, "seed")
        if seed.numel() != n_q:
            raise ValueError("seed must hold one bound per query")
        seed_ptr = seed.data_ptr()
    cap, hoff, hdoc = _impact_cap(dense_cap, hi_off, hi_doc, n_dense)
    pimp = _posting_impacts(post_imp, post_doc)
    with torch.cuda.device(dev):
        check(lib.ragb_bm25_score_topk(_ptr(term_off), _ptr(post_doc), _ptr(post_tf), _ptr(norm), _ptr(idf),
                                       idf.shape[0], k1, dt, stride, dterms, n_dense, imp, maximp, cap, hoff, hdoc, pimp,
                                       _ptr(q_terms), _ptr(q_off), n_q,
                                       max_query_terms, n_docs, id_base, k, seed_ptr, _ptr(score), _ptr(ids), _ptr(ws),
                                       ws.numel(), _stream()))
    return score, ids


@bm25_score_topk.register_fake
def _(term_off, post_doc, post_tf, norm, idf, k1, dense_tf, dense_terms, dense_imp, dense_maximp, q_terms, q_off,
      max_query_terms, id_base, k, seed, dense_cap, hi_off, hi_doc, post_imp):
    n_q = q_off.shape[0] - 1
    return norm.new_empty((n_q, k)), norm.new_empty((n_q, k), dtype=torch.int32)


def bm25_stripe_count(n_queries: int, n_docs: int) -> int:
    """Stripes ``bm25_score_part`` cuts the documents of a shard into for a batch of n_queries (ragb200.h)."""
    return int(lib.ragb_bm25_stripe_count(n_queries, n_docs))


def bm25_workspace(n_queries: int, n_docs: int, k: int, device) -> Tensor:
    return _workspace(lib.ragb_bm25_topk_workspace_bytes(n_queries, n_docs, k), device)


@torch.library.custom_op(f"{NS}::bm25_score_part", mutates_args=("workspace",), device_types="cuda")
def bm25_score_part(term_off: Tensor, post_doc: Tensor, post_tf: Tensor, norm: Tensor, idf: Tensor, k1: float,
                    dense_tf: Tensor, dense_terms: Tensor, dense_imp: Tensor, dense_maximp: Tensor, q_terms: Tensor,
                    q_off: Tensor, max_query_terms: int, id_base: int, k: int, seed: Tensor, dense_cap: Tensor,
                    hi_off: Tensor, hi_doc: Tensor, post_imp: Tensor, stripe_begin: int,
                    stripe_end: int, min_smem_bytes: int, workspace: Tensor) -> None:
    """Score the stripes [stripe_begin, stripe_end) of the staged BM25 search into ``workspace`` (ragb200.h)."""
    term_off, post_doc, post_tf, norm, idf, q_terms, q_off = _bm25_args(term_off, post_doc, post_tf, norm, idf,
                                                                        q_terms, q_off)
    n_q, n_docs, dev = q_off.shape[0] - 1, norm.shape[0], norm.device
    dt, stride, dterms, n_dense = _dense_table(dense_tf, dense_terms, n_docs)
    imp, maximp = None, None
    if n_dense and dense_imp.numel() and dense_maximp.numel():
        imp = _need(dense_imp, torch.float16, "dense_imp").data_ptr()
        maximp = _need(dense_maximp, torch.float32, "dense_maximp").data_ptr()
    seed_ptr = _need(seed, torch.float32, "seed").data_ptr() if seed.numel() else None
    cap, hoff, hdoc = _impact_cap(dense_cap, hi_off, hi_doc, n_dense)
    pimp = _posting_impacts(post_imp, post_doc)
    with torch.cuda.device(dev):
        check(lib.ragb_bm25_score_part(_ptr(term_off), _ptr(post_doc), _ptr(post_tf), _ptr(norm), _ptr(idf), idf.shape[0], k1,
                                       dt, stride, dterms, n_dense, imp, maximp, cap, hoff, hdoc, pimp, _ptr(q_terms), _ptr(q_off), n_q,
                                       max_query_terms, n_docs, id_base, k, seed_ptr, stripe_begin, stripe_end,
                                       min_smem_bytes, _ptr(workspace), workspace.numel(), _stream()))


@torch.library.custom_op(f"{NS}::bm25_score_finish", mutates_args=(), device_types="cuda")
def bm25_score_finish(n_queries: int, n_docs: int, k: int, workspace: Tensor) -> Tuple[Tensor, Tensor]:
    """Merge all stripes of a staged BM25 search: -> (score [B,k], ids [B,k])."""
    dev = workspace.device
    score = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
    ids = torch.empty((n_queries, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.ragb_bm25_score_finish(n_queries, n_docs, k, _ptr(score), _ptr(ids), _ptr(workspace), workspace.numel(),
                                         _stream()))
    return score, ids


@bm25_score_finish.register_fake
def _(n_queries, n_docs, k, workspace):
    return workspace.new_empty((n_queries, k), dtype=torch.float32), workspace.new_empty((n_queries, k), dtype=torch.int32)


@torch.library.custom_op(f"{NS}::bm25_seed", mutates_args=(), device_types="cuda")
def bm25_seed(term_off: Tensor, post_doc: Tensor, post_tf: Tensor, norm: Tensor, idf: Tensor, k1: float,
              dense_tf: Tensor, dense_terms: Tensor, q_terms: Tensor, q_off: Tensor, max_query_terms: int,
              k: int) -> Tensor:
    """Proven lower bounds [B] of every query's k-th best BM25 score over this shard (0 = none); see ragb200.h."""
    term_off, post_doc, post_tf, norm, idf, q_terms, q_off = _bm25_args(term_off, post_doc, post_tf, norm, idf,
                                                                        q_terms, q_off)
    n_q, n_docs, dev = q_off.shape[0] - 1, norm.shape[0], norm.device
    out = torch.empty(n_q, dtype=torch.float32, device=dev)
    dt, stride, dterms, n_dense = _dense_table(dense_tf, dense_terms, n_docs)
    with torch.cuda.device(dev):
        check(lib.ragb_bm25_seed(_ptr(term_off), _ptr(post_doc), _ptr(post_tf), _ptr(norm), _ptr(idf), idf.shape[0], k1,
                                 dt, stride, dterms, n_dense, _ptr(q_terms), _ptr(q_off), n_q, max_query_terms, n_docs, k,
                                 _ptr(out), _stream()))
    return out


@bm25_seed.register_fake
def _(term_off, post_doc, post_tf, norm, idf, k1, dense_tf, dense_terms, q_terms, q_off, max_query_terms, k):
    return norm.new_empty((q_off.shape[0] - 1,))


@torch.library.custom_op(f"{NS}::bm25_scores", mutates_args=(), device_types="cuda")
def bm25_scores(term_off: Tensor, post_doc: Tensor, post_tf: Tensor, norm: Tensor, idf: Tensor, k1: float,
                dense_tf: Tensor, dense_terms: Tensor, q_terms: Tensor, q_off: Tensor,
                max_query_terms: int) -> Tensor:
    term_off, post_doc, post_tf, norm, idf, q_terms, q_off = _bm25_args(term_off, post_doc, post_tf, norm, idf,
                                                                        q_terms, q_off)
    n_q, n_docs, dev = q_off.shape[0] - 1, norm.shape[0], norm.device
    out = torch.empty((n_q, n_docs), dtype=torch.float32, device=dev)
    dt, stride, dterms, n_dense = _dense_table(dense_tf, dense_terms, n_docs)
    with torch.cuda.device(dev):
        check(lib.ragb_bm25_scores(_ptr(term_off), _ptr(post_doc), _ptr(post_tf), _ptr(norm), _ptr(idf),
                                   idf.shape[0], k1, dt, stride, dterms, n_dense, _ptr(q_terms), _ptr(q_off), n_q,
                                   max_query_terms, n_docs, _ptr(out), n_docs, 0, _stream()))
    return out


@bm25_scores.register_fake
def _(term_off, post_doc, post_tf, norm, idf, k1, dense_tf, dense_terms, q_terms, q_off, max_query_terms):
    return norm.new_empty((q_off.shape[0] - 1, norm.shape[0]))


# ------------------------------------------------------------------------------------------
# dense
# ------------------------------------------------------------------------------------------
def _dense_args(passages: Tensor, queries: Tensor):
    passages = _need(passages, torch.bfloat16, "passages")
    queries = _need(queries, torch.bfloat16, "queries")
    if passages.dim() != 2 or queries.dim() != 2 or passages.shape[1] != queries.shape[1]:
        raise ValueError("passages [N, dim] and queries [B, dim] must share dim")
    return passages, queries


@torch.library.custom_op(f"{NS}::dense_gemv_topk", mutates_args=(), device_types="cuda")
def dense_gemv_topk(passages: Tensor, queries: Tensor, k: int, id_base: int) -> Tuple[Tensor, Tensor]:
    passages, queries = _dense_args(passages, queries)
    n_q, dev = queries.shape[0], passages.device
    score = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    ws = _workspace(lib.ragb_dense_gemv_workspace_bytes(n_q, k), dev)
    with torch.cuda.device(dev):
        check(lib.ragb_dense_gemv_topk(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries), n_q, k,
                                       id_base, _ptr(score), _ptr(ids), _ptr(ws), ws.numel(), _stream()))
    return score, ids


@dense_gemv_topk.register_fake
def _(passages, queries, k, id_base):
    n_q = queries.shape[0]
    return (queries.new_empty((n_q, k), dtype=torch.float32), queries.new_empty((n_q, k), dtype=torch.int32))


@torch.library.custom_op(f"{NS}::dense_mma_topk", mutates_args=(), device_types="cuda")
def dense_mma_topk(passages: Tensor, queries: Tensor, k: int, id_base: int, variant: int) -> Tuple[Tensor, Tensor]:
    passages, queries = _dense_args(passages, queries)
    n_q, dev = queries.shape[0], passages.device
    score = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    ws = _workspace(lib.ragb_dense_mma_workspace_bytes(n_q, k), dev)
    with torch.cuda.device(dev):
        check(lib.ragb_dense_mma_topk(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries), n_q, k,
                                      id_base, variant, _ptr(score), _ptr(ids), _ptr(ws), ws.numel(), _stream()))
    return score, ids


@dense_mma_topk.register_fake
def _(passages, queries, k, id_base, variant):
    n_q = queries.shape[0]
    return (queries.new_empty((n_q, k), dtype=torch.float32), queries.new_empty((n_q, k), dtype=torch.int32))


@torch.library.custom_op(f"{NS}::dense_mma_topk_min", mutates_args=(), device_types="cuda")
def dense_mma_topk_min(passages: Tensor, queries: Tensor, k: int, id_base: int, variant: int) -> Tuple[Tensor, Tensor, Tensor]:
    """``dense_mma_topk`` plus the smallest score of every query over the shard (a proven lower bound of it)."""
    passages, queries = _dense_args(passages, queries)
    n_q, dev = queries.shape[0], passages.device
    score = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    lowest = torch.full((n_q,), float("inf"), dtype=torch.float32, device=dev)
    ws = _workspace(lib.ragb_dense_mma_workspace_bytes(n_q, k), dev)
    with torch.cuda.device(dev):
        check(lib.ragb_dense_mma_topk_min(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries), n_q, k, id_base,
                                          variant, _ptr(score), _ptr(ids), _ptr(lowest), _ptr(ws), ws.numel(), _stream()))
    return score, ids, lowest


@dense_mma_topk_min.register_fake
def _(passages, queries, k, id_base, variant):
    n_q = queries.shape[0]
    return (queries.new_empty((n_q, k), dtype=torch.float32), queries.new_empty((n_q, k), dtype=torch.int32),
            queries.new_empty((n_q,), dtype=torch.float32))


@torch.library.custom_op(f"{NS}::dense_mma_sample", mutates_args=(), device_types="cuda")
def dense_mma_sample(passages: Tensor, queries: Tensor, k: int, id_base: int, variant: int) -> Tuple[Tensor, Tensor]:
    """Phase 1 of the seeded tcgen05 search (ragb200.h): -> (thr float32 [B], workspace uint8).  ``thr`` may be raised to
    the maximum over all shards before it is handed to ``dense_mma_seeded`` together with the SAME workspace."""
    passages, queries = _dense_args(passages, queries)
    n_q, dev = queries.shape[0], passages.device
    thr = torch.empty(n_q, dtype=torch.float32, device=dev)
    ws = _workspace(lib.ragb_dense_mma_workspace_bytes(n_q, k), dev)
    with torch.cuda.device(dev):
        check(lib.ragb_dense_mma_sample(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries), n_q, k,
                                        id_base, variant, _ptr(thr), _ptr(ws), ws.numel(), _stream()))
    return thr, ws


@dense_mma_sample.register_fake
def _(passages, queries, k, id_base, variant):
    return queries.new_empty((queries.shape[0],), dtype=torch.float32), queries.new_empty((16,), dtype=torch.uint8)


@torch.library.custom_op(f"{NS}::dense_mma_seeded", mutates_args=("workspace",), device_types="cuda")
def dense_mma_seeded(passages: Tensor, queries: Tensor, k: int, id_base: int, variant: int, thr: Tensor,
                     workspace: Tensor) -> Tuple[Tensor, Tensor]:
    """Phase 2: the rest of the shard with every candidate list seeded by ``thr``, merged with phase 1."""
    passages, queries = _dense_args(passages, queries)
    thr = _need(thr, torch.float32, "thr")
    n_q, dev = queries.shape[0], passages.device
    if thr.numel() != n_q or workspace.dtype != torch.uint8 or not workspace.is_contiguous():
        raise ValueError("thr must hold one bound per query and workspace must be the tensor dense_mma_sample returned")
    score = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.ragb_dense_mma_seeded(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries), n_q, k,
                                        id_base, variant, _ptr(thr), _ptr(score), _ptr(ids), _ptr(workspace),
                                        workspace.numel(), _stream()))
    return score, ids


@dense_mma_seeded.register_fake
def _(passages, queries, k, id_base, variant, thr, workspace):
    n_q = queries.shape[0]
    return (queries.new_empty((n_q, k), dtype=torch.float32), queries.new_empty((n_q, k), dtype=torch.int32))


SCORE_TILE = 256   # passages per tile of the tiled score matrix (ragb200.h, ragb_bm25_scores tiled = 1)


@torch.library.custom_op(f"{NS}::bm25_scores_tiled", mutates_args=("out",), device_types="cuda")
def bm25_scores_tiled(term_off: Tensor, post_doc: Tensor, post_tf: Tensor, norm: Tensor, idf: Tensor, k1: float,
                      dense_tf: Tensor, dense_terms: Tensor, q_terms: Tensor, q_off: Tensor,
                      max_query_terms: int, out: Tensor) -> None:
    """get_scores for a batch written into a caller-owned tiled matrix out[ceil(n_docs / 256), rows >= B, 256] fp32:
    the layout the full-fusion GEMM epilogue reads.  Columns of the last tile beyond n_docs are left untouched."""
    term_off, post_doc, post_tf, norm, idf, q_terms, q_off = _bm25_args(term_off, post_doc, post_tf, norm, idf,
                                                                        q_terms, q_off)
    n_q, n_docs, dev = q_off.shape[0] - 1, norm.shape[0], norm.device
    tiles = -(-n_docs // SCORE_TILE)
    if (out.dtype != torch.float32 or out.dim() != 3 or not out.is_contiguous() or out.shape[0] < tiles
            or out.shape[1] < n_q or out.shape[2] != SCORE_TILE):
        raise ValueError(f"out must be a contiguous float32 [>= {tiles}, >= {n_q}, {SCORE_TILE}] tensor")
    dt, stride, dterms, n_dense = _dense_table(dense_tf, dense_terms, n_docs)
    with torch.cuda.device(dev):
        check(lib.ragb_bm25_scores(_ptr(term_off), _ptr(post_doc), _ptr(post_tf), _ptr(norm), _ptr(idf),
                                   idf.shape[0], k1, dt, stride, dterms, n_dense, _ptr(q_terms), _ptr(q_off), n_q,
                                   max_query_terms, n_docs, _ptr(out), out.shape[1], 1, _stream()))


@torch.library.custom_op(f"{NS}::dense_mma_fused_topk", mutates_args=("counters",), device_types="cuda")
def dense_mma_fused_topk(passages: Tensor, queries: Tensor, bm25: Tensor, w1: Tensor, b1: Tensor, w2: Tensor,
                         b2: Tensor, stats: Tensor, gate_bounds: Tensor, b_cap: float, d_hi: float, k: int, id_base: int,
                         counters: Tensor) -> Tuple[Tensor, Tensor]:
    """Full-fusion top-k inside the tcgen05 epilogue (ragb200.h).

    bm25: the tiled fp32 matrix [ceil(N / 256), rows >= B, 256] ``bm25_scores_tiled`` filled; gate_bounds: int32
    [n_b, n_d], two bf16 (lo, hi) bounds of the gate per cell (``router.full_fusion_bounds``); counters: int64 [2]
    or empty.
    """
    passages, queries = _dense_args(passages, queries)
    n_q, dev = queries.shape[0], passages.device
    tiles = -(-passages.shape[0] // SCORE_TILE)
    if (bm25.dtype != torch.float32 or bm25.dim() != 3 or not bm25.is_contiguous() or bm25.shape[0] < tiles
            or bm25.shape[1] < n_q or bm25.shape[2] != SCORE_TILE):
        raise ValueError(f"bm25 must be a contiguous float32 [>= {tiles}, >= {n_q}, {SCORE_TILE}] tiled score matrix")
    w1, b1, w2, b2, stats = (_need(t, torch.float32, n) for t, n in
                             ((w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2"), (stats, "stats")))
    gate_bounds = _need(gate_bounds, torch.int32, "gate_bounds")
    if gate_bounds.dim() != 2:
        raise ValueError("gate_bounds must be [n_b, n_d]")
    score = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    ws = _workspace(lib.ragb_dense_mma_workspace_bytes(n_q, k), dev)
    cnt = _ptr(counters) if counters.numel() >= 2 else None
    with torch.cuda.device(dev):
        check(lib.ragb_dense_mma_fused_topk(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries), n_q, k,
                                            id_base, _ptr(bm25), bm25.shape[1], _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2),
                                            _ptr(stats), b1.shape[0], _ptr(gate_bounds), gate_bounds.shape[0],
                                            gate_bounds.shape[1],
                                            b_cap, d_hi, _ptr(score), _ptr(ids), cnt, _ptr(ws), ws.numel(), _stream()))
    return score, ids


@dense_mma_fused_topk.register_fake
def _(passages, queries, bm25, w1, b1, w2, b2, stats, gate_bounds, b_cap, d_hi, k, id_base, counters):
    n_q = queries.shape[0]
    return (queries.new_empty((n_q, k), dtype=torch.float32), queries.new_empty((n_q, k), dtype=torch.int32))


@torch.library.custom_op(f"{NS}::bm25_score_docs", mutates_args=(), device_types="cuda")
def bm25_score_docs(term_off: Tensor, post_doc: Tensor, post_tf: Tensor, norm: Tensor, idf: Tensor, k1: float,
                    dense_tf: Tensor, dense_terms: Tensor, q_terms: Tensor, q_off: Tensor, max_query_terms: int,
                    id_base: int, cand_ids: Tensor) -> Tensor:
    """Exact BM25 scores [B, C] of the candidate passages cand_ids int32 [B, C] (global ids, -1 = none -> 0);
    bit-identical to what ``bm25_score_topk`` computes for the same documents."""
    term_off, post_doc, post_tf, norm, idf, q_terms, q_off = _bm25_args(term_off, post_doc, post_tf, norm, idf,
                                                                        q_terms, q_off)
    cand_ids = _need(cand_ids, torch.int32, "cand_ids")
    n_q, n_docs, dev = q_off.shape[0] - 1, norm.shape[0], norm.device
    if cand_ids.dim() != 2 or cand_ids.shape[0] != n_q:
        raise ValueError("cand_ids must be [B, C]")
    out = torch.empty(cand_ids.shape, dtype=torch.float32, device=dev)
    dt, stride, dterms, n_dense = _dense_table(dense_tf, dense_terms, n_docs)
    with torch.cuda.device(dev):
        check(lib.ragb_bm25_score_docs(_ptr(term_off), _ptr(post_doc), _ptr(post_tf), _ptr(norm), _ptr(idf), idf.shape[0], k1,
                                       dt, stride, dterms, n_dense, _ptr(q_terms), _ptr(q_off), n_q, max_query_terms, n_docs,
                                       id_base, _ptr(cand_ids), cand_ids.shape[1], _ptr(out), _stream()))
    return out


@bm25_score_docs.register_fake
def _(term_off, post_doc, post_tf, norm, idf, k1, dense_tf, dense_terms, q_terms, q_off, max_query_terms, id_base, cand_ids):
    return norm.new_empty(cand_ids.shape)


@torch.library.custom_op(f"{NS}::dense_score_docs", mutates_args=(), device_types="cuda")
def dense_score_docs(passages: Tensor, queries: Tensor, id_base: int, cand_ids: Tensor) -> Tensor:
    """Inner products [B, C] of each query with its candidate passages cand_ids int32 [B, C] (-1 = none -> 0)."""
    passages, queries = _dense_args(passages, queries)
    cand_ids = _need(cand_ids, torch.int32, "cand_ids")
    if cand_ids.dim() != 2 or cand_ids.shape[0] != queries.shape[0]:
        raise ValueError("cand_ids must be [B, C]")
    out = torch.empty(cand_ids.shape, dtype=torch.float32, device=passages.device)
    with torch.cuda.device(passages.device):
        check(lib.ragb_dense_score_docs(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries), queries.shape[0],
                                        id_base, _ptr(cand_ids), cand_ids.shape[1], _ptr(out), _stream()))
    return out


@dense_score_docs.register_fake
def _(passages, queries, id_base, cand_ids):
    return queries.new_empty(cand_ids.shape, dtype=torch.float32)


@torch.library.custom_op(f"{NS}::dense_scores", mutates_args=(), device_types="cuda")
def dense_scores(passages: Tensor, queries: Tensor) -> Tensor:
    passages, queries = _dense_args(passages, queries)
    out = torch.empty((queries.shape[0], passages.shape[0]), dtype=torch.float32, device=passages.device)
    with torch.cuda.device(passages.device):
        check(lib.ragb_dense_scores(_ptr(passages), passages.shape[0], passages.shape[1], _ptr(queries),
                                    queries.shape[0], _ptr(out), _stream()))
    return out


@dense_scores.register_fake
def _(passages, queries):
    return queries.new_empty((queries.shape[0], passages.shape[0]), dtype=torch.float32)


# ------------------------------------------------------------------------------------------
# selection / fusion
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::topk_rows", mutates_args=(), device_types="cuda")
def topk_rows(scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    scores = _need(scores, torch.float32, "scores")
    if scores.dim() != 2:
        raise ValueError("scores must be [rows, cols]")
    n_rows, n_cols = scores.shape
    val = torch.empty((n_rows, k), dtype=torch.float32, device=scores.device)
    idx = torch.empty((n_rows, k), dtype=torch.int32, device=scores.device)
    ws = _workspace(lib.ragb_topk_rows_workspace_bytes(n_rows, n_cols, k), scores.device)
    with torch.cuda.device(scores.device):
        check(lib.ragb_topk_rows(_ptr(scores), n_rows, n_cols, k, _ptr(val), _ptr(idx), _ptr(ws), ws.numel(),
                                 _stream()))
    return val, idx


@topk_rows.register_fake
def _(scores, k):
    return scores.new_empty((scores.shape[0], k)), scores.new_empty((scores.shape[0], k), dtype=torch.int32)


@torch.library.custom_op(f"{NS}::topk_merge", mutates_args=(), device_types="cuda")
def topk_merge(scores: Tensor, ids: Tensor, k_out: int) -> Tuple[Tensor, Tensor]:
    """scores / ids [B, n_lists, k_in] -> [B, k_out]"""
    scores = _need(scores, torch.float32, "scores")
    ids = _need(ids, torch.int32, "ids")
    if scores.dim() != 3 or scores.shape != ids.shape:
        raise ValueError("scores and ids must both be [B, n_lists, k_in]")
    n_q, n_lists, k_in = scores.shape
    val = torch.empty((n_q, k_out), dtype=torch.float32, device=scores.device)
    out = torch.empty((n_q, k_out), dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        check(lib.ragb_topk_merge(_ptr(scores), _ptr(ids), n_q, n_lists, k_in, k_out, _ptr(val), _ptr(out),
                                  _stream()))
    return val, out


@topk_merge.register_fake
def _(scores, ids, k_out):
    return scores.new_empty((scores.shape[0], k_out)), ids.new_empty((scores.shape[0], k_out))


@torch.library.custom_op(f"{NS}::topk_merge_gathered", mutates_args=(), device_types="cuda")
def topk_merge_gathered(gathered: Tensor, side: int, k_out: int) -> Tuple[Tensor, Tensor]:
    """Merge one side of the all-gathered exchange buffer in place: ``gathered`` int32 [G, B, 2, n_sides * pool] holds, per
    rank and query, the fp32 score bit patterns (index 0) and the ids (index 1) of n_sides pools next to each other;
    -> the best k_out of pool ``side`` over all ranks, (score [B, k_out], ids [B, k_out])."""
    gathered = _need(gathered, torch.int32, "gathered")
    if gathered.dim() != 4 or gathered.shape[2] != 2:
        raise ValueError("gathered must be [G, B, 2, n_sides * pool]")
    n_ranks, n_q, _, width = gathered.shape
    pool = k_out
    if width % pool or not (0 <= side < width // pool):
        raise ValueError("the last dimension must be a whole number of pools of k_out entries")
    val = torch.empty((n_q, k_out), dtype=torch.float32, device=gathered.device)
    out = torch.empty((n_q, k_out), dtype=torch.int32, device=gathered.device)
    base = gathered.data_ptr() + 4 * side * pool
    with torch.cuda.device(gathered.device):
        check(lib.ragb_topk_merge_strided(base, base + 4 * width, n_q, n_ranks, pool, 2 * width, n_q * 2 * width, k_out,
                                          _ptr(val), _ptr(out), _stream()))
    return val, out


@topk_merge_gathered.register_fake
def _(gathered, side, k_out):
    n_q = gathered.shape[1]
    return gathered.new_empty((n_q, k_out), dtype=torch.float32), gathered.new_empty((n_q, k_out))


@torch.library.custom_op(f"{NS}::hybrid_fuse_topk", mutates_args=(), device_types="cuda")
def hybrid_fuse_topk(bm25_score: Tensor, bm25_id: Tensor, dense_score: Tensor, dense_id: Tensor,
                     k: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Pools [B, pool] each -> (ids, bm25, dense, hybrid) [B, k]."""
    bm25_score = _need(bm25_score, torch.float32, "bm25_score")
    dense_score = _need(dense_score, torch.float32, "dense_score")
    bm25_id = _need(bm25_id, torch.int32, "bm25_id")
    dense_id = _need(dense_id, torch.int32, "dense_id")
    if not (bm25_score.shape == bm25_id.shape == dense_score.shape == dense_id.shape) or bm25_score.dim() != 2:
        raise ValueError("the two pools must be [B, pool] tensors of equal shape")
    n_q, pool = bm25_score.shape
    dev = bm25_score.device
    ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    ob = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    od = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    oh = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.ragb_hybrid_fuse_topk(_ptr(bm25_score), _ptr(bm25_id), _ptr(dense_score), _ptr(dense_id), n_q, pool,
                                        k, _ptr(ids), _ptr(ob), _ptr(od), _ptr(oh), _stream()))
    return ids, ob, od, oh


@hybrid_fuse_topk.register_fake
def _(bm25_score, bm25_id, dense_score, dense_id, k):
    n_q = bm25_score.shape[0]
    f = bm25_score.new_empty((n_q, k))
    return bm25_id.new_empty((n_q, k)), f, f.clone(), f.clone()


@torch.library.custom_op(f"{NS}::retrieval_uncertainty", mutates_args=(), device_types="cuda")
def retrieval_uncertainty(scores: Tensor, ids: Tensor, lam: float) -> Tensor:
    """U = std(top-k scores) + lam * (1 - |s_1 - s_k|) per ranked list [B, k] (ids < 0 are padding)."""
    scores = _need(scores, torch.float32, "scores")
    ids = _need(ids, torch.int32, "ids")
    if scores.dim() != 2 or scores.shape != ids.shape:
        raise ValueError("scores and ids must both be [B, k]")
    out = torch.empty(scores.shape[0], dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        check(lib.ragb_retrieval_uncertainty(_ptr(scores), _ptr(ids), scores.shape[0], scores.shape[1], lam, _ptr(out),
                                             _stream()))
    return out


@retrieval_uncertainty.register_fake
def _(scores, ids, lam):
    return scores.new_empty((scores.shape[0],))


# ------------------------------------------------------------------------------------------
# router
# ------------------------------------------------------------------------------------------
def _router_args(bm25, dense, w1, b1, w2, b2, stats):
    bm25 = _need(bm25, torch.float32, "bm25")
    dense = _need(dense, torch.float32, "dense")
    if bm25.dim() != 2 or bm25.shape != dense.shape:
        raise ValueError("bm25 and dense must both be [B, P]")
    w1 = _need(w1, torch.float32, "w1")
    b1 = _need(b1, torch.float32, "b1")
    w2 = _need(w2, torch.float32, "w2").reshape(-1)
    b2 = _need(b2, torch.float32, "b2").reshape(-1)
    stats = _need(stats, torch.float32, "stats").reshape(-1)
    hidden = b1.shape[0]
    if w1.shape != (hidden, 3) or w2.shape[0] != hidden or b2.shape[0] != 1 or stats.shape[0] != 4:
        raise ValueError("router weights must be w1 [H,3], b1 [H], w2 [H] or [1,H], b2 [1], stats [4]")
    return bm25, dense, w1, b1, w2, b2, stats, hidden


@torch.library.custom_op(f"{NS}::router_forward", mutates_args=(), device_types="cuda")
def router_forward(bm25: Tensor, dense: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, stats: Tensor,
                   norm_mode: int) -> Tuple[Tensor, Tensor]:
    """-> (gate [B,P], fused [B,P])"""
    bm25, dense, w1, b1, w2, b2, stats, hidden = _router_args(bm25, dense, w1, b1, w2, b2, stats)
    n_rows, n_cand = bm25.shape
    gate = torch.empty_like(bm25)
    fused = torch.empty_like(bm25)
    scratch = _workspace(lib.ragb_router_scratch_bytes(n_rows, norm_mode), bm25.device)
    with torch.cuda.device(bm25.device):
        check(lib.ragb_router_forward(_ptr(bm25), _ptr(dense), n_rows, n_cand, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2),
                                      _ptr(stats), hidden, norm_mode, _ptr(gate), _ptr(fused), _ptr(scratch),
                                      _stream()))
    return gate, fused


@router_forward.register_fake
def _(bm25, dense, w1, b1, w2, b2, stats, norm_mode):
    return torch.empty_like(bm25), torch.empty_like(bm25)


@torch.library.custom_op(f"{NS}::router_mc_dropout", mutates_args=(), device_types="cuda")
def router_mc_dropout(bm25: Tensor, dense: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, stats: Tensor,
                      norm_mode: int, n_samples: int, p_drop: float, seed: int, offset: int, mask_layout: int,
                      dump: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> mean_gate, std_gate, mean_fused, std_fused [B,P]; variance [B]; consensus [B];
    mask_dump uint8 [T, B*P, H] and gate_dump [T,B,P] (empty unless ``dump``)."""
    bm25, dense, w1, b1, w2, b2, stats, hidden = _router_args(bm25, dense, w1, b1, w2, b2, stats)
    n_q, n_cand = bm25.shape
    dev = bm25.device
    outs = [torch.empty_like(bm25) for _ in range(4)]
    variance = torch.empty(n_q, dtype=torch.float32, device=dev)
    consensus = torch.empty(n_q, dtype=torch.int32, device=dev)
    if dump:
        mask = torch.empty((n_samples, n_q * n_cand, hidden), dtype=torch.uint8, device=dev)
        gates = torch.empty((n_samples, n_q, n_cand), dtype=torch.float32, device=dev)
    else:
        mask = torch.empty(0, dtype=torch.uint8, device=dev)
        gates = torch.empty(0, dtype=torch.float32, device=dev)
    scratch = _workspace(lib.ragb_router_scratch_bytes(n_q, norm_mode), dev)
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    with torch.cuda.device(dev):
        check(lib.ragb_router_mc_dropout(_ptr(bm25), _ptr(dense), n_q, n_cand, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2),
                                         _ptr(stats), hidden, norm_mode, n_samples, p_drop, seed & 0xFFFFFFFFFFFFFFFF, offset, mask_layout,
                                         sm_count, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]),
                                         _ptr(variance), _ptr(consensus), _ptr(mask) if dump else None,
                                         _ptr(gates) if dump else None, _ptr(scratch), _stream()))
    return outs[0], outs[1], outs[2], outs[3], variance, consensus, mask, gates


@router_mc_dropout.register_fake
def _(bm25, dense, w1, b1, w2, b2, stats, norm_mode, n_samples, p_drop, seed, offset, mask_layout, dump):
    n_q, n_cand = bm25.shape
    e = [torch.empty_like(bm25) for _ in range(4)]
    hidden = b1.shape[0]
    mask = bm25.new_empty((n_samples, n_q * n_cand, hidden) if dump else (0,), dtype=torch.uint8)
    gates = bm25.new_empty((n_samples, n_q, n_cand) if dump else (0,))
    return (e[0], e[1], e[2], e[3], bm25.new_empty((n_q,)), bm25.new_empty((n_q,), dtype=torch.int32), mask, gates)


def launch_count() -> int:
    """Number of kernels libragb200 has launched in this process (bench.py's gpu_launches)."""
    return _lib.launch_count()
