"""CPU restatement of dense scoring, top-k and pool fusion.  TEST INFRASTRUCTURE ONLY.

Follows
  * ``DenseIndex.search``                  ``rag_uq/streaming_index.py:338-370``
    (ChromaDB cosine-space HNSW, ``score = 1 - distance``).  The oracle is the
    EXACT cosine the approximate index is trying to reach, evaluated on the
    bf16-rounded unit rows the product stores, accumulated in float64.
  * ``HybridRetriever.hybrid_search``      ``rag_uq/streaming_index.py:464-523``
  * ``HybridRetriever.get_scores_for_router`` ``rag_uq/streaming_index.py:525-557``

PINNED (fusion): ``tests/golden/make_retrieval_golden.py`` runs the LIVE
``rag_uq.streaming_index.HybridRetriever`` from /root/reference with its two
``*_search`` methods replaced by fixed pools and stores what ``hybrid_search`` and
``get_scores_for_router`` return (``tests/golden/retrieval_golden.json``);
``tests/test_oracle_cpu.py`` checks the two functions below against it bit for bit
(and against the live class on random pools whenever /root/reference is mounted).
PARITY UNPINNED (dense scoring only): chromadb is not installable in this image,
so the exact cosine is anchored on its definition, not on ChromaDB output.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def dense_scores(passages_f32: np.ndarray, queries_f32: np.ndarray) -> np.ndarray:
    """[B, N] float64 inner products of already bf16-rounded (float32-held) rows."""
    return queries_f32.astype(np.float64) @ passages_f32.astype(np.float64).T


def topk_desc(scores: np.ndarray, k: int, positive_only: bool = False) -> List[List[Tuple[int, float]]]:
    """Row-wise top-k with the framework's tie rule: score desc, then index asc.

    ``positive_only`` reproduces ``if scores[idx] > 0`` (streaming_index.py:176).
    """
    out = []
    idx = np.arange(scores.shape[1])
    for row in scores:
        order = np.lexsort((idx, -row))[:k]
        out.append([(int(i), float(row[i])) for i in order if (row[i] > 0 or not positive_only)])
    return out


def hybrid_search(bm25_pool: Sequence[Tuple[int, float]], dense_pool: Sequence[Tuple[int, float]],
                  top_k: int = 10, known_ids=None) -> List[Tuple[int, float, float, float]]:
    """Pool fusion, streaming_index.py:484-523.  Returns (id, bm25, dense, hybrid) rows.

    Union of the two pools, ids the retriever holds no document for are skipped
    BEFORE the maxima are taken (:494-496; ``known_ids`` = container of the ids it
    holds, None = all), a score missing from one pool is 0.0 (:498-499), each
    score divided by the pool-union maximum - ``max(...) or 1`` (:513-514) -
    averaged (:517-519), sorted descending (:521), cut to top_k (:523).  Tie order
    in the reference depends on set iteration order; we fix it to id ascending.
    """
    b: Dict[int, float] = dict(bm25_pool)
    d: Dict[int, float] = dict(dense_pool)
    ids = sorted(i for i in (set(b) | set(d)) if known_ids is None or i in known_ids)
    if not ids:
        return []
    rows = [(i, b.get(i, 0.0), d.get(i, 0.0)) for i in ids]
    max_b = max(r[1] for r in rows) or 1
    max_d = max(r[2] for r in rows) or 1
    fused = [(i, sb, sd, (sb / max_b + sd / max_d) / 2) for (i, sb, sd) in rows]
    fused.sort(key=lambda r: (-r[3], r[0]))
    return fused[:top_k]


def scores_for_router(bm25_pool, dense_pool, num_passages: int = 20, known_ids=None):
    """streaming_index.py:537-557: hybrid_search(top_k=num_passages) split into aligned
    lists and padded with 0.0 / id -1 (the reference pads ids with "").  The caller hands
    in pools of (at most) 50: get_scores_for_router does not forward a pool size (:537)."""
    rows = hybrid_search(bm25_pool, dense_pool, top_k=num_passages, known_ids=known_ids)
    bm = [r[1] for r in rows]
    de = [r[2] for r in rows]
    ids = [r[0] for r in rows]
    while len(bm) < num_passages:
        bm.append(0.0)
        de.append(0.0)
        ids.append(-1)
    return bm, de, ids


def merge_topk(parts: Sequence[Sequence[Tuple[int, float]]], k: int) -> List[Tuple[int, float]]:
    """Merge per-shard candidate lists into one global top-k (score desc, id asc)."""
    flat = [c for part in parts for c in part]
    flat.sort(key=lambda c: (-c[1], c[0]))
    return flat[:k]


def retrieval_uncertainty(scores: Sequence[float], lam: float = 1.0) -> float:
    """docs/uncertainty_theory.md:48-56: U = std(s_top-k) + lambda * (1 - |s_1 - s_k|)."""
    if len(scores) == 0:
        return lam
    arr = np.asarray(scores, dtype=np.float64)
    return float(arr.std() + lam * (1.0 - abs(arr.max() - arr.min())))
