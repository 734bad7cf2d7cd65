"""CPU restatement of the RetrievalRouter gate, rerank and MC-Dropout.  TEST INFRASTRUCTURE ONLY.

Follows
  * ``RetrievalRouter._normalize_scores`` ``rag_uq/router.py:100-138``
  * ``RetrievalRouter.forward``           ``rag_uq/router.py:140-177``
  * ``RetrievalRouter.hybrid_rerank``     ``rag_uq/router.py:179-202``
  * Dropout placement                     ``rag_uq/router.py:73-83`` (Linear, ReLU, Dropout, Linear, Sigmoid)
  * MC aggregation math                   ``rag_uq/confidence.py:195-202, 258-264``

PINNED: ``tests/golden/make_golden.py`` runs the live ``rag_uq.router`` module
from /root/reference and stores inputs + outputs in
``tests/golden/router_golden.npz``; ``tests/test_oracle_cpu.py`` checks every
function here against those vectors (and against the live module when
/root/reference is present).

Weights are passed as a plain dict with the reference's state-dict keys
(``scorer.0.weight`` [H,3], ``scorer.0.bias`` [H], ``scorer.3.weight`` [1,H],
``scorer.3.bias`` [1], ``bm25_mean``, ``bm25_std``, ``dense_mean``, ``dense_std``).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-6  # router.py:112


def normalize(bm25: torch.Tensor, dense: torch.Tensor, state: Dict[str, torch.Tensor],
              stats_initialized: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """router.py:130-136 (the inference branches; the EMA update at :114-128 is training-only)."""
    if stats_initialized:
        bn = (bm25 - state["bm25_mean"]) / (state["bm25_std"] + EPS)
        dn = (dense - state["dense_mean"]) / (state["dense_std"] + EPS)
    else:
        # whole-tensor statistics, torch.std is the unbiased (n-1) estimator
        bn = (bm25 - bm25.mean()) / (bm25.std() + EPS)
        dn = (dense - dense.mean()) / (dense.std() + EPS)
    return bn, dn


def gate(bm25: torch.Tensor, dense: torch.Tensor, state: Dict[str, torch.Tensor],
         stats_initialized: bool, keep_mask: Optional[torch.Tensor] = None,
         p_drop: float = 0.1) -> torch.Tensor:
    """router.py:158-177.  ``keep_mask`` ([B*P, H] of 0/1) injects a dropout mask
    after the ReLU exactly where nn.Dropout sits (:78): h * mask / (1 - p)."""
    bn, dn = normalize(bm25, dense, state, stats_initialized)
    feats = torch.stack([bn, dn, dn - bn], dim=-1).view(-1, 3)
    hidden = torch.relu(F.linear(feats, state["scorer.0.weight"], state["scorer.0.bias"]))
    if keep_mask is not None:
        hidden = hidden * keep_mask / (1.0 - p_drop)
    out = torch.sigmoid(F.linear(hidden, state["scorer.3.weight"], state["scorer.3.bias"]))
    return out.view(bm25.shape)


def fused(bm25: torch.Tensor, dense: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """router.py:199 - on the RAW scores."""
    return w * dense + (1 - w) * bm25


def hybrid_rerank(bm25, dense, state, stats_initialized, top_k: int = 10):
    """router.py:196-202; ties resolved score desc / index asc (torch.topk leaves it open)."""
    h = fused(bm25, dense, gate(bm25, dense, state, stats_initialized))
    k = min(top_k, h.size(-1))
    order = torch.argsort(-h.double(), dim=-1, stable=True)[:, :k]
    return torch.gather(h, 1, order), order


def mc_dropout(bm25, dense, state, stats_initialized, keep_masks: torch.Tensor, p_drop: float = 0.1):
    """T stochastic passes (keep_masks [T, B*P, H]) aggregated as confidence.py does.

    Per candidate: mean / population-std (ddof = 0, like ``distances.std()`` at
    confidence.py:200) of the gate and of the fused score.  Per query: the
    T gate vectors play the role of the T answer embeddings - centroid, L2
    distance of every sample to it, ``variance = distances.std()``
    (:196-200), ``uncertainty = min(1, variance / 2)``, ``confidence = 1 - u``
    (:258-264), consensus = argmin distance (:248-249).
    """
    ws = torch.stack([gate(bm25, dense, state, stats_initialized, keep_masks[t], p_drop)
                      for t in range(keep_masks.shape[0])])          # [T, B, P]
    hs = torch.stack([fused(bm25, dense, w) for w in ws])
    out = {
        "mean_w": ws.mean(0), "std_w": ws.std(0, unbiased=False),
        "mean_h": hs.mean(0), "std_h": hs.std(0, unbiased=False),
    }
    centroid = ws.mean(0, keepdim=True)
    dist = torch.linalg.norm(ws - centroid, dim=-1)                   # [T, B]
    variance = dist.std(0, unbiased=False)
    unc = torch.clamp(variance / 2.0, max=1.0)
    out.update(variance=variance, uncertainty=unc, confidence=1.0 - unc, consensus=dist.argmin(0))
    return out
