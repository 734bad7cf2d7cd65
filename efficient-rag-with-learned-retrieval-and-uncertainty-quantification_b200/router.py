"""Drop-in ``RetrievalRouter`` whose inference path runs in libragb200's kernels.

Mirrors rag_uq/router.py:34-232: same ``RouterConfig``, same parameter / buffer names
(``scorer.0.weight`` [H,3], ``scorer.0.bias`` [H], ``scorer.3.weight`` [1,H], ``scorer.3.bias``
[1], ``bm25_mean``, ``bm25_std``, ``dense_mean``, ``dense_std``) so ``load_state_dict`` of a
reference checkpoint (router.py:499-517, loaded at experiments/run_evaluation.py:128-131) works
unchanged, same ``forward`` / ``hybrid_rerank`` / ``get_routing_decision`` signatures and return
types.  As in the reference, ``stats_initialized`` is a plain attribute that is NOT part of the
state dict: straight after ``load_state_dict`` the gate normalises with the statistics of the
call itself (router.py:133-136) until the caller sets ``stats_initialized = True``.

Training (``ApproxNDCGLoss`` / ``RouterTrainer``, router.py:235-561) is outside the hot path:
the kernels produce no autograd graph.  ``forward`` raises if a gradient is requested.

New on top of the reference: ``mc_dropout`` - T stochastic passes with the Dropout of
router.py:78 active, drawn in-kernel from Philox4x32-10, aggregated with the arithmetic of
``MCDropoutConfidence`` (rag_uq/confidence.py:195-202, 258-264).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .confidence import ConfidenceResult, RouterUncertainty

logger = logging.getLogger(__name__)


@dataclass
class RouterConfig:
    """rag_uq/router.py:34-41, field for field (checkpoints pickle this class by value)."""
    hidden_dim: int = 64
    dropout: float = 0.1
    temperature: float = 1.0
    num_layers: int = 2
    use_batch_norm: bool = False


NORM_BATCH, NORM_RUNNING, NORM_PER_QUERY = 0, 1, 2


class RetrievalRouter(nn.Module):
    def __init__(self, config: Optional[Any] = None):
        super().__init__()
        self.config = config or RouterConfig()
        hidden = int(self.config.hidden_dim)
        if int(self.config.num_layers) != 2 or bool(getattr(self.config, "use_batch_norm", False)):
            raise ValueError("rag_uq_b200.RetrievalRouter implements num_layers=2, use_batch_norm=False "
                             "(the reference defaults, router.py:40-41); there is no fallback for other shapes")
        if hidden % 4 or not (4 <= hidden <= 128):
            raise ValueError("hidden_dim must be a multiple of 4 in [4, 128]")
        # identical module tree => identical state-dict keys and identical default initialisation
        self.scorer = nn.Sequential(
            nn.Linear(3, hidden), nn.ReLU(), nn.Dropout(float(self.config.dropout)), nn.Linear(hidden, 1), nn.Sigmoid()
        )
        self.register_buffer("bm25_mean", torch.tensor(0.0))
        self.register_buffer("bm25_std", torch.tensor(1.0))
        self.register_buffer("dense_mean", torch.tensor(0.0))
        self.register_buffer("dense_std", torch.tensor(1.0))
        self.stats_initialized = False
        logger.info(f"Initialized RetrievalRouter with {self._count_params()} parameters")

    def _count_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    # ------------------------------------------------------------------------------------
    def _weights(self):
        lin1, lin2 = self.scorer[0], self.scorer[3]
        stats = torch.stack([self.bm25_mean, self.bm25_std, self.dense_mean, self.dense_std]).to(torch.float32)
        return (lin1.weight.detach(), lin1.bias.detach(), lin2.weight.detach().reshape(-1), lin2.bias.detach(), stats)

    def _check(self, bm25_scores: torch.Tensor, dense_scores: torch.Tensor):
        if not bm25_scores.is_cuda:
            raise NotImplementedError("rag_uq_b200.RetrievalRouter runs on CUDA (sm_100) tensors only; "
                                      "move the module and its inputs to the GPU (there is no CPU path)")
        if self.scorer[0].weight.device != bm25_scores.device:
            raise RuntimeError("router parameters and scores live on different devices")
        if torch.is_grad_enabled() and (bm25_scores.requires_grad or dense_scores.requires_grad or
                                        (self.training and any(p.requires_grad for p in self.parameters()))):
            raise NotImplementedError("the B200 router kernels are inference-only (no autograd); train with the "
                                      "reference RouterTrainer and load its checkpoint, or call under torch.no_grad()")
        return bm25_scores.to(torch.float32), dense_scores.to(torch.float32)

    def _update_running_stats(self, bm25_scores, dense_scores):
        """EMA of router.py:114-128 (training-mode side effect; tiny host-driven reductions)."""
        eps, momentum = 1e-6, 0.1
        with torch.no_grad():
            self.bm25_mean = (1 - momentum) * self.bm25_mean + momentum * bm25_scores.mean()
            self.bm25_std = (1 - momentum) * self.bm25_std + momentum * (bm25_scores.std() + eps)
            self.dense_mean = (1 - momentum) * self.dense_mean + momentum * dense_scores.mean()
            self.dense_std = (1 - momentum) * self.dense_std + momentum * (dense_scores.std() + eps)
            self.stats_initialized = True

    def forward(self, bm25_scores: torch.Tensor, dense_scores: torch.Tensor, update_stats: bool = True,
                per_query_stats: bool = False) -> torch.Tensor:
        """Gate weights [B, P] in (0, 1); 0 favours BM25, 1 favours dense (router.py:140-177).

        ``per_query_stats`` (extension): when running statistics are not armed, normalise every
        row with its own mean / std - what the reference's evaluation loop gets by calling the
        router once per query with a [1, P] tensor.
        """
        bm25_scores, dense_scores = self._check(bm25_scores, dense_scores)
        if update_stats and self.training:
            self._update_running_stats(bm25_scores, dense_scores)
        mode = NORM_RUNNING if self.stats_initialized else (NORM_PER_QUERY if per_query_stats else NORM_BATCH)
        w1, b1, w2, b2, stats = self._weights()
        if self.training and self.config.dropout > 0:
            seed, offset = _philox_state(bm25_scores.device, bm25_scores.numel() * w1.shape[0])
            out = ops.router_mc_dropout(bm25_scores, dense_scores, w1, b1, w2, b2, stats, mode, 1,
                                        float(self.config.dropout), seed, offset, 1, True)
            return out[7][0]
        gate, _ = ops.router_forward(bm25_scores, dense_scores, w1, b1, w2, b2, stats, mode)
        return gate

    def hybrid_rerank(self, bm25_scores: torch.Tensor, dense_scores: torch.Tensor,
                      top_k: int = 10, per_query_stats: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """(top-k fused scores, int64 indices), fused = w*dense + (1-w)*bm25 (router.py:179-202).

        ``per_query_stats`` (extension, as in ``forward``): without armed running statistics every row is
        normalised with its own mean / std - the result of calling the reference once per query with [1, P]
        tensors, which is what its evaluation loop does (experiments/run_evaluation.py:165-184).  The default
        normalises over the whole call, exactly like the reference does for a [B, P] call.
        In train mode the reference's ``forward(update_stats=False)`` applies Dropout (router.py:196): so does this.
        """
        bm25_scores, dense_scores = self._check(bm25_scores, dense_scores)
        mode = NORM_RUNNING if self.stats_initialized else (NORM_PER_QUERY if per_query_stats else NORM_BATCH)
        w1, b1, w2, b2, stats = self._weights()
        if self.training and self.config.dropout > 0:
            gate = self.forward(bm25_scores, dense_scores, update_stats=False, per_query_stats=per_query_stats)
            fused = (gate * dense_scores + (1 - gate) * bm25_scores).contiguous()
        else:
            _, fused = ops.router_forward(bm25_scores, dense_scores, w1, b1, w2, b2, stats, mode)
        k = min(top_k, fused.size(-1))
        values, indices = ops.topk_rows(fused, k)
        return torch.return_types.topk((values, indices.to(torch.int64)))

    def get_routing_decision(self, bm25_scores: torch.Tensor, dense_scores: torch.Tensor,
                             threshold: float = 0.5) -> Dict[str, Any]:
        """router.py:204-232."""
        with torch.no_grad():
            was_training = self.training
            self.eval()
            weights = self.forward(bm25_scores, dense_scores, update_stats=False)
            self.train(was_training)
            return {
                "avg_dense_weight": weights.mean().item(),
                "weight_std": weights.std().item(),
                "dense_preferred_ratio": (weights > threshold).float().mean().item(),
                "bm25_preferred_ratio": (weights <= threshold).float().mean().item(),
                "routing_weights": weights.cpu().numpy(),
            }

    def full_fusion_table(self, b_cap: float, d_hi: float, n_b: int = 128, n_d: int = 64) -> torch.Tensor:
        """Device copy of ``full_fusion_bounds`` for this router, cached until a weight or statistic changes."""
        tensors = list(self._weights()[:4]) + [self.bm25_mean, self.bm25_std, self.dense_mean, self.dense_std]
        key = (float(b_cap), float(d_hi), int(n_b), int(n_d), tuple((t.data_ptr(), t._version) for t in tensors))
        cache = self.__dict__.setdefault("_ff_cache", {})
        if key not in cache:
            # entries of other (b_cap, d_hi) grids stay (the fallback of the threshold search alternates between a few);
            # entries of older weights or statistics go
            for stale in [c for c in cache if c[-1] != key[-1]] + (list(cache)[:1] if len(cache) >= 16 else []):
                cache.pop(stale, None)
            w1, b1, w2, b2, stats = (t.detach().cpu().numpy() for t in self._weights())
            cache[key] = torch.from_numpy(full_fusion_bounds(w1, b1, w2, b2, stats, b_cap, d_hi, n_b, n_d)).to(
                self.scorer[0].weight.device)
        return cache[key]

    def full_fusion_envelope(self, b_cap: float, d_hi: float, n_b: int = 128, n_d: int = 64) -> torch.Tensor:
        """float32 [n_b, n_d] on the router's device: E[ib, id] >= fused(b', d') for EVERY b' <= (ib + 1) b_cap / n_b and
        every d' in column id (``full_fusion_envelope`` below); cached like the bound table."""
        tensors = list(self._weights()[:4]) + [self.bm25_mean, self.bm25_std, self.dense_mean, self.dense_std]
        key = ("env", float(b_cap), float(d_hi), int(n_b), int(n_d), tuple((t.data_ptr(), t._version) for t in tensors))
        cache = self.__dict__.setdefault("_env_cache", {})
        if key not in cache:
            for stale in [c for c in cache if c[-1] != key[-1]] + (list(cache)[:1] if len(cache) >= 16 else []):
                cache.pop(stale, None)
            w1, b1, w2, b2, stats = (t.detach().cpu().numpy() for t in self._weights())
            cache[key] = torch.from_numpy(full_fusion_envelope(w1, b1, w2, b2, stats, b_cap, d_hi, n_b, n_d)).to(
                self.scorer[0].weight.device)
        return cache[key]

    # ------------------------------------------------------------------------------------
    def mc_dropout(self, bm25_scores: torch.Tensor, dense_scores: torch.Tensor, n_samples: int = 30,
                   seed: Optional[int] = None, offset: int = 0, torch_layout: bool = False,
                   per_query_stats: bool = False, return_samples: bool = False) -> RouterUncertainty:
        """T stochastic gate evaluations per candidate with the router's Dropout active.

        ``seed`` None draws (seed, offset) from torch's CUDA generator and advances it the way T
        ``F.dropout`` calls would.  ``torch_layout`` maps elements to Philox counters exactly as
        torch's fused CUDA dropout does for the [B*P, H] hidden tensor, so sample t equals the
        t-th ``router(..., update_stats=False)`` call of a reference module in train mode.
        """
        bm25_scores, dense_scores = self._check(bm25_scores, dense_scores)
        mode = NORM_RUNNING if self.stats_initialized else (NORM_PER_QUERY if per_query_stats else NORM_BATCH)
        w1, b1, w2, b2, stats = self._weights()
        if seed is None:
            seed, offset = _philox_state(bm25_scores.device, bm25_scores.numel() * w1.shape[0], n_samples)
        out = ops.router_mc_dropout(bm25_scores, dense_scores, w1, b1, w2, b2, stats, mode, int(n_samples),
                                    float(self.config.dropout), int(seed), int(offset), 1 if torch_layout else 0,
                                    bool(return_samples))
        return RouterUncertainty(mean_gate=out[0], std_gate=out[1], mean_fused=out[2], std_fused=out[3],
                                 variance=out[4], consensus=out[5].to(torch.int64), n_samples=int(n_samples),
                                 masks=out[6] if return_samples else None, gates=out[7] if return_samples else None)


def full_fusion_gate_range(w1, b1, w2, b2, stats, b_cap: float, d_hi: float, n_b: int = 128, n_d: int = 64,
                           gate_slack: float = 1e-4) -> Tuple[np.ndarray, np.ndarray]:
    """Proven bounds (lo, hi), float64 [n_b, n_d], of the running-statistics gate on a grid of cells.

    lo[ib, id] <= gate(b, d) <= hi[ib, id] for every bm25 score b in [ib, ib+1] * b_cap / n_b and every
    dense score d in -d_hi + [id, id+1] * 2 d_hi / n_d (router.py:130-132, 158-177).

    How: the gate's pre-activation z(b, d) = b2 + sum_j w2_j * relu(A_j b + D_j d + C_j) is continuous
    and piecewise linear, so over a rectangular cell it takes its extremes at a vertex of the line
    arrangement: a cell corner, a crossing of a unit's zero line with a cell edge, or a crossing of two
    zero lines inside the cell.  All of them are enumerated in float64 and sigma is monotone.
    ``gate_slack`` covers the fp32 evaluation in the kernel.
    """
    w1 = np.asarray(w1, dtype=np.float64).reshape(-1, 3)
    b1 = np.asarray(b1, dtype=np.float64).reshape(-1)
    w2 = np.asarray(w2, dtype=np.float64).reshape(-1)
    b2 = float(np.asarray(b2, dtype=np.float64).reshape(-1)[0])
    st = np.asarray(stats, dtype=np.float32).reshape(4)
    sb = float(np.float32(st[1] + np.float32(1e-6)))   # the kernel divides by float32(std + 1e-6)
    sd = float(np.float32(st[3] + np.float32(1e-6)))
    mb, md = float(st[0]), float(st[2])
    lo, hi = np.zeros((n_b, n_d)), np.ones((n_b, n_d))
    if not (np.isfinite(w1).all() and np.isfinite(b1).all() and np.isfinite(w2).all() and np.isfinite(b2)
            and np.isfinite([sb, sd, mb, md]).all() and sb != 0.0 and sd != 0.0 and b_cap > 0 and d_hi > 0):
        return lo, hi   # [0, 1] is always valid (bound = max(b, d): prunes little)
    A = (w1[:, 0] - w1[:, 2]) / sb
    D = (w1[:, 1] + w1[:, 2]) / sd
    Cc = b1 - (w1[:, 0] - w1[:, 2]) * mb / sb - (w1[:, 1] + w1[:, 2]) * md / sd
    wb, wd = b_cap / n_b, 2.0 * d_hi / n_d
    b_edges = np.arange(n_b + 1, dtype=np.float64) * wb
    d_edges = -d_hi + np.arange(n_d + 1, dtype=np.float64) * wd

    def z_of(b, d):
        pre = np.multiply.outer(b, A) + np.multiply.outer(d, D) + Cc
        return b2 + np.maximum(pre, 0.0) @ w2

    zmin = np.full((n_b, n_d), np.inf)
    zmax = np.full((n_b, n_d), -np.inf)

    def offer_point(b, d, z):
        """A vertex belongs to every cell whose closed rectangle contains it (within 1e-6 cells of a grid line: both sides)."""
        pb, pd = b / wb, (d + d_hi) / wd
        for ib in (np.floor(pb - 1e-6).astype(np.int64), np.floor(pb + 1e-6).astype(np.int64)):
            for idx in (np.floor(pd - 1e-6).astype(np.int64), np.floor(pd + 1e-6).astype(np.int64)):
                ok = (ib >= 0) & (ib < n_b) & (idx >= 0) & (idx < n_d)
                np.minimum.at(zmin, (ib[ok], idx[ok]), z[ok])
                np.maximum.at(zmax, (ib[ok], idx[ok]), z[ok])

    gb, gd = np.meshgrid(b_edges, d_edges, indexing="ij")                     # cell corners
    offer_point(gb.reshape(-1), gd.reshape(-1), z_of(gb.reshape(-1), gd.reshape(-1)))
    pts_b, pts_d = [], []
    with np.errstate(divide="ignore", invalid="ignore"):
        dc = -(np.multiply.outer(b_edges, A) + Cc) / D                        # zero lines x vertical grid lines
        pts_b.append(np.broadcast_to(b_edges[:, None], dc.shape).reshape(-1))
        pts_d.append(dc.reshape(-1))
        bc = -(np.multiply.outer(d_edges, D) + Cc) / A                        # zero lines x horizontal grid lines
        pts_b.append(bc.reshape(-1))
        pts_d.append(np.broadcast_to(d_edges[:, None], bc.shape).reshape(-1))
        i, j = np.triu_indices(A.shape[0], 1)                                 # crossings of two zero lines
        det = A[i] * D[j] - A[j] * D[i]
        pts_b.append((-Cc[i] * D[j] + Cc[j] * D[i]) / det)
        pts_d.append((-A[i] * Cc[j] + A[j] * Cc[i]) / det)
    bx, dx = np.concatenate(pts_b), np.concatenate(pts_d)
    ok = np.isfinite(bx) & np.isfinite(dx) & (bx >= 0) & (bx <= b_cap) & (dx >= -d_hi) & (dx <= d_hi)
    if ok.any():
        offer_point(bx[ok], dx[ok], z_of(bx[ok], dx[ok]))
    sig = lambda x: 1.0 / (1.0 + np.exp(-np.clip(x, -700.0, 700.0)))  # noqa: E731
    return np.clip(sig(zmin) - gate_slack, 0.0, 1.0), np.clip(sig(zmax) + gate_slack, 0.0, 1.0)


def _bf16_bits(x: np.ndarray, up: bool) -> np.ndarray:
    """bfloat16 bit patterns of non-negative float64 values, rounded down (up=False) or up (up=True)."""
    f = x.astype(np.float32)
    f = np.where(f.astype(np.float64) > x, np.nextafter(f, np.float32(-1.0)), f) if not up else \
        np.where(f.astype(np.float64) < x, np.nextafter(f, np.float32(2.0)), f)
    bits = f.astype(np.float32).view(np.uint32)
    if up:
        bits = bits + np.where(bits & 0xFFFF, np.uint32(0x10000), np.uint32(0))
    return (bits >> 16).astype(np.uint32)


def full_fusion_bounds(w1, b1, w2, b2, stats, b_cap: float, d_hi: float, n_b: int = 128, n_d: int = 64) -> np.ndarray:
    """The gate-bound table ragb_dense_mma_fused_topk takes (include/ragb200.h): int32 [n_b, n_d], bits 0-15 = bf16
    lower bound (rounded down), bits 16-31 = bf16 upper bound (rounded up) of the gate on each cell; the last bm25
    row is (0, 1) because it also receives bm25 >= b_cap and bm25 < 0.  The table only prunes gate evaluations
    (fused = b + g (d - b) <= b + (d <= b ? lo : hi) (d - b)), it never changes results."""
    lo, hi = full_fusion_gate_range(w1, b1, w2, b2, stats, b_cap, d_hi, n_b, n_d)
    lo[n_b - 1, :] = 0.0
    hi[n_b - 1, :] = 1.0
    packed = (_bf16_bits(hi, True) << 16) | _bf16_bits(lo, False)
    return packed.astype(np.uint32).view(np.int32)


def full_fusion_envelope(w1, b1, w2, b2, stats, b_cap: float, d_hi: float, n_b: int = 128, n_d: int = 64) -> np.ndarray:
    """Upper envelope, monotone in the BM25 score, of the fused score over the grid of ``full_fusion_gate_range``.

    fused(b, d) = b + g (d - b) with lo <= g <= hi on a cell is linear in g and bilinear in (b, d), so over a cell it is
    at most the maximum over the four corners and the two gate bounds; E[ib, id] is the running maximum of that over all
    rows ib' <= ib of column id: a proven bound on the fused score of ANY passage whose BM25 score is at most the upper
    edge of row ib and whose dense score lies in column id.  The maximum of E[ib, id_lo .. id_hi] therefore bounds every
    passage with bm25 <= b and d_lo <= dense <= d_hi - the stopping rule of the threshold-algorithm search
    (engine.full_fusion_topk).  The dense range has to be two-sided: the gate moves with the dense score, and
    (1 - g) * bm25 evaluated at dense scores no passage has would inflate the bound.  float32 [n_b, n_d], rounded up."""
    lo, hi = full_fusion_gate_range(w1, b1, w2, b2, stats, b_cap, d_hi, n_b, n_d)
    b_edges = np.arange(n_b + 1, dtype=np.float64) * (b_cap / n_b)
    d_edges = -d_hi + np.arange(n_d + 1, dtype=np.float64) * (2.0 * d_hi / n_d)
    cell = np.full((n_b, n_d), -np.inf)
    for bb in (b_edges[:-1], b_edges[1:]):
        for dd in (d_edges[:-1], d_edges[1:]):
            diff = dd[None, :] - bb[:, None]
            for g in (lo, hi):
                cell = np.maximum(cell, bb[:, None] + g * diff)
    env = np.maximum.accumulate(cell, axis=0)
    out = env.astype(np.float32)
    return np.where(out.astype(np.float64) < env, np.nextafter(out, np.float32(np.inf)), out).astype(np.float32)


def torch_dropout_increment(n_elements: int, sm_count: int) -> int:
    """Philox offset consumed by one torch fused-dropout call over n float elements."""
    block = 256
    grid = min(sm_count * (2048 // block), (n_elements + block - 1) // block)
    return ((n_elements - 1) // (block * grid * 4) + 1) * 4


def _philox_state(device, n_elements: int, n_calls: int = 1) -> Tuple[int, int]:
    """(seed, offset) of torch's CUDA generator, advanced as n_calls dropout launches would."""
    index = torch.device(device).index
    gen = torch.cuda.default_generators[torch.cuda.current_device() if index is None else index]
    seed, offset = int(gen.initial_seed()), int(gen.get_offset())
    inc = torch_dropout_increment(n_elements, torch.cuda.get_device_properties(device).multi_processor_count)
    gen.set_offset(offset + inc * n_calls)
    # torch.seed() draws 64 random bits: about half of those seeds are >= 2^63.  The op schema carries a signed
    # int64, so such a seed travels as its two's-complement image and is widened back to the same 64 bits at the
    # C boundary (ops.router_mc_dropout) - masking it to 63 bits would change the Philox key.
    return seed - (1 << 64) if seed >= (1 << 63) else seed, offset


__all__ = ["RouterConfig", "RetrievalRouter", "ConfidenceResult", "RouterUncertainty"]
