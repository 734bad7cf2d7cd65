"""Tensor-level hybrid retrieval over one row shard of the corpus, optionally one of G shards.

This is the batched form of ``HybridRetriever.hybrid_search`` (rag_uq/streaming_index.py:464-523)
and ``get_scores_for_router`` (:525-557): BM25 pool + dense pool -> union / max-normalised
average -> top-k.  The string-level drop-in classes in ``retrieval.py`` sit on top of it.

Sharding: passages are independent, so GPU g owns global rows
[g*N/G, (g+1)*N/G): its slice of the embedding matrix and a document-partitioned inverted index
built with GLOBAL statistics (df all-reduced once at build time).  Queries are replicated.  Per
batch there is exactly one exchange: each rank all-gathers its local [B, pool] candidate lists
(fp32 score + int32 global id; NCCL over NVLink) and every rank merges G*pool candidates per
query.  Pool fusion needs the GLOBAL pools (max-normalisation and the 0.0-for-missing rule), so
it runs after the merge.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib, ops
from .sparse import SparseShard


# sharded search: raise the pruning bounds of both kernels to their maximum over the shards before the kernels run
EXCHANGE_BOUNDS = os.environ.get("RAGB_EXCHANGE_BOUNDS", "1") != "0"
SLICE_SEEDS = os.environ.get("RAGB_SLICE_SEEDS", "1") != "0"   # sharded search: every shard seeds 1/world of the batch
OVERLAP_FRACTION = float(os.environ.get("RAGB_OVERLAP_FRACTION", "1.0"))   # share of the BM25 stripes scored beside the GEMM
OVERLAP_SMEM_PAD = int(os.environ.get("RAGB_OVERLAP_SMEM_KB", "0")) * 1024     # shared memory per BM25 block while it is


def shard_rows(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank``; the first ``n_rows % world`` ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def global_bm25_statistics(df_local: Tensor, n_local: int, len_local: int, group=None):
    """Sum document frequencies, document count and total length over the shards."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return df_local, n_local, len_local
    df = df_local.clone()
    scal = torch.tensor([n_local, len_local], dtype=torch.int64, device=df_local.device)
    dist.all_reduce(df, group=group)
    dist.all_reduce(scal, group=group)
    return df, int(scal[0]), int(scal[1])


def gather_candidates(score: Tensor, ids: Tensor, group=None) -> Tuple[Tensor, Tensor]:
    """All-gather local candidate lists [B, k] into [B, G, k] (rank-major inside a query)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return score.unsqueeze(1), ids.unsqueeze(1)
    b, k = score.shape
    # ONE collective: fp32 scores travel as their int32 bit pattern next to the int32 ids ([B, 2, k] per rank; the
    # message is latency-bound - 800 KB per rank at B = 1024, two pools of 50 - so one launch instead of two matters)
    if score.dtype != torch.float32 or ids.dtype != torch.int32:
        raise TypeError("gather_candidates expects float32 scores and int32 ids")
    packed = torch.stack([score.contiguous().view(torch.int32), ids.contiguous()], dim=1)
    flat = torch.empty((world * b, 2, k), dtype=torch.int32, device=score.device)   # rank-major concatenation
    dist.all_gather_into_tensor(flat, packed, group=group)
    gathered = flat.view(world, b, 2, k)
    return (gathered[:, :, 0].permute(1, 0, 2).contiguous().view(torch.float32),
            gathered[:, :, 1].permute(1, 0, 2).contiguous())


def exchange_pools(bs: Tensor, bi: Tensor, ds: Tensor, di: Tensor, group=None):
    """The one exchange of a sharded hybrid search: all-gather the two local pools of every rank ([B, 2, 2 * pool] int32 per
    rank: fp32 score bit patterns next to the ids, ONE collective) and merge each side over the ranks straight out of the
    gathered buffer (ragb_topk_merge_strided: no transposing copies).  -> global (bs, bi, ds, di), identical on every rank."""
    world = dist.get_world_size(group)
    n_q, pool = bs.shape
    packed = torch.empty((n_q, 2, 2 * pool), dtype=torch.int32, device=bs.device)
    packed[:, 0, :pool], packed[:, 0, pool:] = bs.view(torch.int32), ds.view(torch.int32)
    packed[:, 1, :pool], packed[:, 1, pool:] = bi, di
    gathered = torch.empty((world, n_q, 2, 2 * pool), dtype=torch.int32, device=bs.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    gbs, gbi = ops.topk_merge_gathered(gathered, 0, pool)
    gds, gdi = ops.topk_merge_gathered(gathered, 1, pool)
    return gbs, gbi, gds, gdi


def seed_slice(n_queries: int, rank: int, world: int) -> Tuple[int, int]:
    """Queries [q0, q1) of a batch that shard ``rank`` of ``world`` seeds before the bound exchange: contiguous slices of
    ceil(n / world) that cover every query exactly once (trailing shards may get none)."""
    per = -(-n_queries // max(1, world))
    return min(n_queries, rank * per), min(n_queries, (rank + 1) * per)


class HybridEngine:
    def __init__(self, sparse: Optional[SparseShard], passages: Optional[Tensor], id_base: int = 0, group=None,
                 mma_variant: int = 3):
        if passages is not None:
            if passages.dtype != torch.bfloat16 or passages.dim() != 2:
                raise TypeError("passages must be a bf16 [rows, dim] tensor")
            if not passages.is_cuda:
                raise NotImplementedError("rag_uq_b200 scores on CUDA (sm_100) only; there is no CPU path")
        self.sparse = sparse
        self.passages = passages
        self.id_base = id_base
        self.group = group
        self.mma_variant = mma_variant
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._side_stream = None

    # ---- single-retriever pools --------------------------------------------------------
    def _merge(self, score: Tensor, ids: Tensor, k: int) -> Tuple[Tensor, Tensor]:
        if self.world == 1:
            return score, ids
        s, i = gather_candidates(score, ids, self.group)
        return ops.topk_merge(s, i, k)

    def bm25_topk(self, q_terms: Tensor, q_off: Tensor, max_terms: int, k: int) -> Tuple[Tensor, Tensor]:
        """BM25Index.search for a batch: (score [B,k], global id [B,k]); score > 0 only, id -1 pads."""
        score, ids = self.sparse.score_topk(q_terms, q_off, max_terms, k)
        return self._merge(score, ids, k)

    def dense_local_topk(self, q_emb: Tensor, k: int) -> Tuple[Tensor, Tensor]:
        if q_emb.shape[0] <= _lib.GEMV_MAX_BATCH:
            return ops.dense_gemv_topk(self.passages, q_emb, k, self.id_base)
        return ops.dense_mma_topk(self.passages, q_emb, k, self.id_base, self.mma_variant)

    def dense_topk(self, q_emb: Tensor, k: int) -> Tuple[Tensor, Tensor]:
        """DenseIndex.search for a batch, exact instead of HNSW."""
        score, ids = self.dense_local_topk(q_emb, k)
        return self._merge(score, ids, k)

    # ---- hybrid ------------------------------------------------------------------------
    def local_pools(self, q_terms: Tensor, q_off: Tensor, max_terms: int, q_emb: Tensor, pool: int,
                    overlap: bool = False, events=None):
        """Local BM25 pool and local dense pool of one batch.

        ``overlap`` runs the two scoring kernels on two streams (dense first so its one-block-per-SM
        grid gets its SMs; needs RAGB_MMA_STAGES=3 so two BM25 blocks fit beside it).  Measured on B200
        at 10M passages x 1024 queries: 47.4 ms overlapped vs 47.8 ms back to back - the spinning
        producer / issuer warps and the epilogue of the tensor-core kernel compete with the issue-bound
        BM25 kernel for the same issue slots - so it is off by default.
        ``events``: optional dict filled with (start, end) CUDA events per kernel, recorded on the stream
        the kernel runs on.
        """
        cur = torch.cuda.current_stream()
        big = q_emb.shape[0] > _lib.GEMV_MAX_BATCH

        def mark():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        # Sharded corpus: both kernels prune against a proven lower bound of the query's k-th best score (the BM25
        # seed; the k-th best of the dense kernel's sampled prefix).  The k-th best of the WHOLE corpus is at least
        # the k-th best of any shard, so the bounds are raised to their maximum over the shards first: one tiny
        # all-reduce ([2, B] floats, latency-bound) in front of the two kernels.  Results do not depend on it.
        exchange = self.world > 1 and big and hasattr(self.sparse, "seed") and EXCHANGE_BOUNDS
        if exchange:
            t0 = mark() if events is not None else None
            # The seed kernel costs the same on a shard of any size (<= 1024 postings per list term, random gathers:
            # 0.17 ms per 1024 queries - 9 % of the BM25 time of a 1.25M-row shard), and ANY shard's bound is valid
            # for the whole corpus.  So each shard seeds only its slice of the batch (q_off is sliced, its offsets stay
            # absolute) and the MAX all-reduce below hands every query the bound of the shard that seeded it.
            n_q = q_off.shape[0] - 1
            q0, q1 = seed_slice(n_q, self.rank, self.world)
            b_seed = torch.zeros(n_q, dtype=torch.float32, device=q_emb.device)
            if SLICE_SEEDS and q1 > q0:
                b_seed[q0:q1] = self.sparse.seed(q_terms, q_off[q0:q1 + 1], max_terms, pool)
            elif not SLICE_SEEDS:
                b_seed = self.sparse.seed(q_terms, q_off, max_terms, pool)
            ta = mark() if events is not None else None
            d_thr, d_ws = ops.dense_mma_sample(self.passages, q_emb, pool, self.id_base, self.mma_variant)
            tb = mark() if events is not None else None
            both = torch.stack([b_seed, d_thr])
            dist.all_reduce(both, op=dist.ReduceOp.MAX, group=self.group)
            t1 = mark() if events is not None else None
            bs, bi = self.sparse.score_topk(q_terms, q_off, max_terms, pool, both[0])
            t2 = mark() if events is not None else None
            ds, di = ops.dense_mma_seeded(self.passages, q_emb, pool, self.id_base, self.mma_variant, both[1], d_ws)
            if events is not None:
                events["bm25_seed"], events["dense_prefix"], events["seed_exchange"] = (t0, ta), (ta, tb), (tb, t1)
                events["bm25"], events["dense"] = (t1, t2), (t2, mark())
            return bs, bi, ds, di

        if not (overlap and big):
            t0 = mark() if events is not None else None
            bs, bi = self.sparse.score_topk(q_terms, q_off, max_terms, pool)
            t1 = mark() if events is not None else None
            ds, di = self.dense_local_topk(q_emb, pool)
            if events is not None:
                events["bm25"], events["dense"] = (t0, t1), (t1, mark())
            return bs, bi, ds, di
        # ---- two streams, co-resident blocks -------------------------------------------------------------------
        # The tensor-core kernel needs one 320-thread block per SM, ~38k registers and (with a 4-stage operand ring)
        # 129 KB of shared memory: what is left of an SM hosts exactly one 8-warp BM25 block if that block's shared
        # memory is padded to ~96 KB (two of them no longer fit).  So the search is staged: the first stripes of the
        # BM25 search run NEXT TO the GEMM (one block per SM, a third of the kernel's usual residency), the remaining
        # stripes after it at full residency, starting from the thresholds the first part has proven.  The dense
        # kernel's stream has the higher priority and its main kernel becomes eligible (event after the sampled
        # prefix) when the padded BM25 part does, so the GEMM blocks are placed first.
        if not isinstance(self.sparse, SparseShard):
            raise NotImplementedError("overlap=True needs a single-segment BM25 shard")
        dev = q_emb.device
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=dev, priority=-1)
        side = self._side_stream
        n_q, n_docs = q_emb.shape[0], self.sparse.n_docs
        stripes = ops.bm25_stripe_count(n_q, n_docs)
        first = min(stripes, max(0, int(round(stripes * OVERLAP_FRACTION))))
        ws = ops.bm25_workspace(n_q, n_docs, pool, dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            d0 = mark() if events is not None else None
            variant = 4 if self.mma_variant == 3 else self.mma_variant      # 4-stage ring: room for a co-resident block
            d_thr, d_ws = ops.dense_mma_sample(self.passages, q_emb, pool, self.id_base, variant)
            prefix_done = torch.cuda.Event()
            prefix_done.record()
            ds, di = ops.dense_mma_seeded(self.passages, q_emb, pool, self.id_base, variant, d_thr, d_ws)
            d1 = mark() if events is not None else None
        b0 = mark() if events is not None else None
        cur.wait_event(prefix_done)
        self.sparse.score_part(q_terms, q_off, max_terms, pool, ws, 0, first, OVERLAP_SMEM_PAD)
        bmid = mark() if events is not None else None
        cur.wait_stream(side)
        if first < stripes:
            self.sparse.score_part(q_terms, q_off, max_terms, pool, ws, first, stripes, 0)
        bs, bi = ops.bm25_score_finish(n_q, n_docs, pool, ws)
        b1 = mark() if events is not None else None
        for t in (ds, di, d_ws):
            t.record_stream(cur)
        ws.record_stream(side)
        if events is not None:
            events["bm25"], events["dense"], events["bm25_beside_dense"] = (b0, b1), (d0, d1), (b0, bmid)
        return bs, bi, ds, di

    def hybrid_topk(self, q_terms: Tensor, q_off: Tensor, max_terms: int, q_emb: Tensor, k: int = 10,
                    pool: int = 50, overlap: bool = False):
        """-> ids int32 [B,k] (-1 pads), bm25 [B,k], dense [B,k], hybrid [B,k]."""
        bs, bi, ds, di = self.local_pools(q_terms, q_off, max_terms, q_emb, pool, overlap)
        if self.world > 1:
            bs, bi, ds, di = exchange_pools(bs, bi, ds, di, self.group)
        return ops.hybrid_fuse_topk(bs, bi, ds, di, k)

    def retrieve_and_rerank(self, q_terms: Tensor, q_off: Tensor, max_terms: int, q_emb: Tensor, router, k: int = 10,
                            pool: int = 50, mc_samples: int = 0, lam: float = 1.0, seed: Optional[int] = None):
        """The per-query loop of experiments/run_evaluation.py:165-196 for a whole batch.

        get_scores_for_router (hybrid top-k, padded) -> router gate -> fused score -> reorder, plus the
        confidence the reference left as a TODO (:194-196): the retrieval uncertainty of
        docs/uncertainty_theory.md:48-56 on the fused ranking and, with ``mc_samples`` > 0, the
        MC-Dropout confidence of the router on these candidates.  Returns a dict of device tensors.
        """
        ids, sb, sd, sh = self.hybrid_topk(q_terms, q_off, max_terms, q_emb, k, pool)
        # The reference loop calls the router once per query with [1, k] tensors (run_evaluation.py:171-174): until
        # running statistics are armed (the state right after load_state_dict) every query is normalised with ITS OWN
        # mean / std.  A [B, k] call would normalise over the whole batch and make a query's ranking depend on its
        # batch mates, so the per-query mode is requested explicitly; with running statistics it changes nothing.
        per_query = not getattr(router, "stats_initialized", False)
        fused, order = router.hybrid_rerank(sb, sd, top_k=k, per_query_stats=per_query)
        ranked = torch.gather(ids, 1, order)
        out = {"ids": ranked, "fused": fused, "bm25": torch.gather(sb, 1, order), "dense": torch.gather(sd, 1, order),
               "hybrid": torch.gather(sh, 1, order),
               "retrieval_uncertainty": ops.retrieval_uncertainty(fused.contiguous(), ranked.contiguous(), lam)}
        if mc_samples > 0:
            unc = router.mc_dropout(sb, sd, n_samples=mc_samples, seed=seed, per_query_stats=per_query)
            out["router_confidence"] = unc.confidence
            out["gate_mean"], out["gate_std"] = unc.mean_gate, unc.std_gate
        return out

    # ---- latency path: the whole small-batch step as ONE CUDA graph --------------------------------
    def graphed_search(self, router, batch: int, max_terms: int, k: int = 10, pool: int = 50) -> "GraphedSearch":
        """Capture ``retrieve_and_rerank`` for ``batch`` <= 8 queries of at most ``max_terms`` tokens into a CUDA graph
        (at batch 1 the step is ~10 launches for a few hundred microseconds of work, so launch latency
        and the serial BM25 -> GEMV order dominate).  Inside the graph the BM25 chain and the GEMV run as two
        parallel branches; replaying it costs one launch.  Single shard only (an NCCL exchange is not captured)."""
        return GraphedSearch(self, router, batch, max_terms, k, pool)

    # ---- full-fusion mode ----------------------------------------------------------------
    def _max_passage_norm(self) -> float:
        """Largest L2 norm of a passage row (computed once, chunked): with the largest query norm it bounds |dense|."""
        if getattr(self, "_max_norm", None) is None:
            best = torch.zeros((), dtype=torch.float32, device=self.passages.device)
            for r0 in range(0, self.passages.shape[0], 1 << 18):
                best = torch.maximum(best, self.passages[r0:r0 + (1 << 18)].float().norm(dim=1).max())
            self._max_norm = float(best)
        return self._max_norm

    def bm25_score_cap(self, q_terms: Tensor, q_off: Tensor) -> float:
        """An upper bound of every BM25 score of the batch: tf (k1+1) / (tf + norm) < k1 + 1 per occurrence."""
        idf = self.sparse.idf
        ok = (q_terms >= 0) & (q_terms < idf.shape[0])
        w = torch.where(ok, idf[q_terms.clamp(0, idf.shape[0] - 1).long()].clamp(min=0.0), idf.new_zeros(()))
        csum = torch.cat([w.new_zeros(1), torch.cumsum(w.double(), 0).float()])
        per_query = csum[q_off[1:].long()] - csum[q_off[:-1].long()]
        k1 = float(getattr(self.sparse, "k1", 1.5))
        return float(per_query.max()) * (k1 + 1.0) if per_query.numel() else 0.0

    def full_fusion_topk(self, q_terms: Tensor, q_off: Tensor, max_terms: int, q_emb: Tensor, router, k: int = 10,
                         query_chunk: Optional[int] = None, fused: Optional[bool] = None, counters: Optional[Tensor] = None,
                         events=None, method: str = "auto", depth: int = 64, info: Optional[dict] = None):
        """Gate evaluated on the true scores of every (query, passage) pair: the k best fused scores per query.

        ``method`` "auto" (default) / "threshold": the threshold-algorithm search below, with the exhaustive epilogue as
        the fallback for queries whose stopping rule does not hold; "exhaustive": the [B, N]-scan described under
        ``_full_fusion_exhaustive`` (``fused`` / ``query_chunk`` / ``counters`` apply to it).
        """
        if method == "exhaustive" or fused is not None or not isinstance(self.sparse, SparseShard) \
                or q_emb.shape[0] <= _lib.GEMV_MAX_BATCH:
            return self._full_fusion_exhaustive(q_terms, q_off, max_terms, q_emb, router, k, query_chunk, fused, counters, events)
        return self._full_fusion_threshold(q_terms, q_off, max_terms, q_emb, router, k, depth, events, info)

    def _full_fusion_threshold(self, q_terms, q_off, max_terms, q_emb, router, k, depth, events, info):
        """Full-fusion as a threshold-algorithm (Fagin) search over the two exact ranked lists - no [B, N] matrix.

        fused(b, d) = b + g(b, d) (d - b) is bounded from above by a function E that is monotone in both scores
        (``router.full_fusion_envelope``).  So: (1) the streaming kernels deliver the ``depth`` best passages of each
        side, exactly, and the dense kernel also the smallest dense score of the shard; (2) every listed passage gets its
        OTHER score (ragb_bm25_score_docs / ragb_dense_score_docs) and its exact gate and fused score; (3) any passage on
        neither list has bm25 <= b_depth and d_min <= dense <= d_depth, hence fused <= max E(b_depth, [d_min, d_depth]):
        if that is below the k-th best fused score found, the top-k is proven complete.
        Queries for which it is not (a gate that ignores both scores' order, a flat score distribution) are re-run
        through the exhaustive epilogue.  Oracle: RetrievalRouter.hybrid_rerank(bm25[B,N], dense[B,N], k)."""
        if not getattr(router, "stats_initialized", False):
            raise ValueError("full-fusion mode needs router.stats_initialized = True (running statistics)")
        n_q, n_local = q_emb.shape[0], self.passages.shape[0]
        if int(self.sparse.id_base) != int(self.id_base):
            raise ValueError("the BM25 shard and the engine number their passages from different bases "
                             f"({self.sparse.id_base} vs {self.id_base}): the two ranked lists could not be joined")
        # two ranked lists of `depth` passages each (at least 2 k: the rule needs the depth-th score clearly below the
        # k-th); 64 measured best at k = 10 on 10M passages (31.6 ms against 32.8 ms at 100, one fallback query either way)
        depth = max(k, min(max(depth, 2 * k), _lib.MMA_MAX_TOPK, n_local))
        e0 = _mark(events)
        bs, bi = self.sparse.score_topk(q_terms, q_off, max_terms, depth)
        e1 = _mark(events)
        ds, di, d_lowest = ops.dense_mma_topk_min(self.passages, q_emb, depth, self.id_base, self.mma_variant)
        e2 = _mark(events)
        # union of the two lists: a passage on both keeps its BM25-list slot
        dup = (di.unsqueeze(2) == bi.unsqueeze(1)).any(dim=2) & (di >= 0)
        cand = torch.cat([bi, torch.where(dup, torch.full_like(di, -1), di)], dim=1).contiguous()
        b_c = self.sparse.score_docs(q_terms, q_off, max_terms, cand)
        d_c = ops.dense_score_docs(self.passages, q_emb, self.id_base, cand)
        w1, b1, w2, b2, stats = router._weights()
        _, fused_c = ops.router_forward(b_c, d_c, w1, b1, w2, b2, stats, 1)
        fused_c = torch.where(cand >= 0, fused_c, torch.full_like(fused_c, float("-inf"))).contiguous()
        kk = min(k, n_local)
        val, idx = ops.topk_rows(fused_c, kk)
        ids = torch.gather(cand, 1, idx.to(torch.int64).clamp(min=0))
        ids = torch.where(idx >= 0, ids, torch.full_like(ids, -1))
        # stopping rule: bound on everything unlisted (an incomplete BM25 list means every other passage scores 0)
        b_cap = self.bm25_score_cap(q_terms, q_off)
        b_cap = float(min(64.0, 2.0 ** max(0, int(b_cap - 1e-9).bit_length()))) if b_cap > 0 else 1.0
        d_hi = self._max_passage_norm() * float(q_emb.float().norm(dim=1).max()) * 1.002 + 1e-3
        d_hi = float(-(-d_hi * 64 // 1) / 64)
        # a fine grid (cells of b_cap / 1024 x d_hi / 128): on a large corpus the best BM25 scores lie close together,
        # so the rule has to separate the depth-th from the k-th score by less than a coarse cell
        env = router.full_fusion_envelope(b_cap, d_hi, 1024, 256)
        n_b, n_d = env.shape
        b_last = torch.where(bi[:, -1] >= 0, bs[:, -1], torch.zeros_like(bs[:, -1]))
        d_last = ds[:, -1] + 2e-6                        # tensor-core vs fp32 summation order
        d_first = d_lowest - 2e-6                        # the smallest dense score any passage of the shard has
        ib = torch.clamp((b_last * (n_b / b_cap)).floor().long(), 0, n_b - 1)
        id_hi = torch.clamp(((d_last + d_hi) * (n_d / (2.0 * d_hi))).floor().long(), 0, n_d - 1)
        id_lo = torch.clamp(((d_first + d_hi) * (n_d / (2.0 * d_hi))).floor().long(), 0, n_d - 1)
        cols = torch.arange(n_d, device=env.device)
        inside = (cols[None, :] >= id_lo[:, None]) & (cols[None, :] <= id_hi[:, None])
        unseen = torch.where(inside, env[ib], torch.full((), float("-inf"), device=env.device)).amax(dim=1)
        kth = val[:, kk - 1]
        proven = (unseen <= kth - 1e-5 * kth.abs() - 1e-6) | (depth >= n_local)
        e3 = _mark(events)
        redo = torch.nonzero(~proven).flatten()
        n_redo = int(redo.numel())
        if info is not None:
            info.update({"method": "threshold", "depth": depth, "fallback_queries": n_redo})
        if n_redo:
            sel = redo.tolist()
            q_off_h = q_off.tolist()
            sub_terms = torch.cat([q_terms[q_off_h[q]:q_off_h[q + 1]] for q in sel]) if any(q_off_h[q + 1] > q_off_h[q] for q in sel) \
                else q_terms[:1]
            lens = torch.tensor([q_off_h[q + 1] - q_off_h[q] for q in sel], dtype=torch.int32, device=q_off.device)
            sub_off = torch.cat([lens.new_zeros(1), torch.cumsum(lens, 0).to(torch.int32)]).contiguous()
            if n_redo <= _lib.GEMV_MAX_BATCH:           # the epilogue kernel wants a real batch: pad by repeating queries
                pad = (_lib.GEMV_MAX_BATCH + 1) - n_redo
                sub_emb = torch.cat([q_emb[redo], q_emb[redo[:1]].expand(pad, -1)]).contiguous()
                sub_terms = torch.cat([sub_terms, sub_terms[:int(lens[0])].repeat(pad)]) if int(lens[0]) else sub_terms
                sub_off = torch.cat([sub_off, sub_off[-1] + (torch.arange(1, pad + 1, device=q_off.device, dtype=torch.int32) * int(lens[0]))])
            else:
                sub_emb = q_emb[redo].contiguous()
            # the batch's own (b_cap, d_hi) grid: valid for any subset of it, and the same cached bound table every step
            fv, fi = self._full_fusion_exhaustive(sub_terms.contiguous(), sub_off.contiguous(), max_terms, sub_emb, router, k,
                                                  None, None, None, None, merge=False, bounds=(b_cap, d_hi))
            val[redo], ids[redo] = fv[:n_redo, :kk], fi[:n_redo, :kk]
        if events is not None:
            events.setdefault("bm25", []).append((e0, e1))
            events.setdefault("dense", []).append((e1, e2))
            events.setdefault("candidates", []).append((e2, e3))
        return self._merge(val, ids, val.shape[1])

    def _full_fusion_exhaustive(self, q_terms: Tensor, q_off: Tensor, max_terms: int, q_emb: Tensor, router, k: int = 10,
                                query_chunk: Optional[int] = None, fused: Optional[bool] = None,
                                counters: Optional[Tensor] = None, events=None, merge: bool = True,
                                bounds: Optional[Tuple[float, float]] = None):
        """Gate evaluated for EVERY (query, passage) pair on the true scores ("full-fusion").

        Oracle: ``RetrievalRouter.hybrid_rerank(bm25_full[B,N], dense_full[B,N], k)`` (router.py:179-202).
        The router must be in running-statistics mode: call-wide statistics over [B, N] would need a
        global reduction first and are refused here.  -> (fused score [B,k], global id int32 [B,k]).

        ``fused`` (default: whenever the batch is large enough for the tensor-core kernel): the BM25
        kernel writes get_scores for a chunk of queries into a tiled [N_local / 256, chunk, 256] fp32 matrix and the
        tcgen05 GEMM's epilogue reads it, evaluates gate and fusion on the accumulator in TMEM and keeps
        the per-query top-k: neither the dense nor the fused score matrix ever exists.  Otherwise the
        un-fused form materialises both matrices per chunk (small shapes, tests).
        ``query_chunk`` defaults to what fits in 60 % of the free device memory.
        """
        if not getattr(router, "stats_initialized", False):
            raise ValueError("full-fusion mode needs router.stats_initialized = True (running statistics)")
        w1, b1, w2, b2, stats = router._weights()
        n_q, n_local = q_emb.shape[0], self.passages.shape[0]
        if fused is None:
            fused = n_q > _lib.GEMV_MAX_BATCH and k <= _lib.MMA_MAX_TOPK and hasattr(self.sparse, "scores_tiled")
        tiles = -(-n_local // ops.SCORE_TILE)
        ld = tiles * ops.SCORE_TILE
        if query_chunk is None:
            per_query = ld * 4 * (1 if fused else 3)
            if n_q * per_query <= (1 << 30):
                query_chunk = n_q        # small (the fallback of the threshold search): no cudaMemGetInfo, which was seen to
            else:                        # take 5-50 ms now and then when called between kernels
                free, _ = torch.cuda.mem_get_info(q_emb.device)
                query_chunk = max(1, min(n_q, int(free * 0.6) // per_query))
            if fused and query_chunk >= 128:
                query_chunk = query_chunk // 128 * 128
        out_s, out_i = [], []
        q_off_host = q_off.tolist()
        if fused:
            if bounds is not None:      # a caller that already bounded a superset of these queries
                b_cap, d_hi = bounds
            else:
                b_cap = self.bm25_score_cap(q_terms, q_off)
                b_cap = float(min(64.0, 2.0 ** max(0, int(b_cap - 1e-9).bit_length()))) if b_cap > 0 else 1.0
                d_hi = self._max_passage_norm() * float(q_emb.float().norm(dim=1).max()) * 1.002 + 1e-3
                d_hi = float(-(-d_hi * 64 // 1) / 64)   # round up to 1/64 so the cached table is reused across batches
            gate_bounds = router.full_fusion_table(b_cap, d_hi)
            # the tiled BM25 matrix of a chunk is kept between calls when it is small (the fallback of the threshold search
            # asks for the same ~360 MB every step; a fresh allocation while the previous step is still in flight
            # occasionally cost a cudaMalloc of tens of milliseconds)
            shape = (tiles, min(query_chunk, n_q), ops.SCORE_TILE)
            keep = getattr(self, "_ff_bm", None)
            if keep is not None and tuple(keep.shape) == shape and keep.device == q_emb.device:
                bm = keep
            else:
                bm = torch.empty(shape, dtype=torch.float32, device=q_emb.device)
                self._ff_bm = bm if bm.numel() * 4 <= (1 << 30) else None
            if counters is None:
                counters = torch.empty(0, dtype=torch.int64, device=q_emb.device)
        for lo in range(0, n_q, query_chunk):
            hi = min(n_q, lo + query_chunk)
            t0, t1 = q_off_host[lo], q_off_host[hi]
            sub_terms = q_terms[t0:t1] if t1 > t0 else q_terms[:1]
            sub_off = (q_off[lo:hi + 1] - t0).contiguous()
            if fused:
                e0 = _mark(events)
                self.sparse.scores_tiled(sub_terms.contiguous(), sub_off, max_terms, bm)
                e1 = _mark(events)
                val, idx = ops.dense_mma_fused_topk(self.passages, q_emb[lo:hi].contiguous(), bm, w1, b1, w2, b2,
                                                    stats, gate_bounds, b_cap, d_hi, min(k, n_local), self.id_base, counters)
                if events is not None:
                    events.setdefault("bm25", []).append((e0, e1))
                    events.setdefault("dense", []).append((e1, _mark(events)))
                out_s.append(val)
                out_i.append(idx)
                continue
            bm_full = self.sparse.scores(sub_terms.contiguous(), sub_off, max_terms)
            de = ops.dense_scores(self.passages, q_emb[lo:hi].contiguous())
            _, fused_scores = ops.router_forward(bm_full, de, w1, b1, w2, b2, stats, 1)
            val, idx = ops.topk_rows(fused_scores, min(k, fused_scores.shape[1]))
            out_s.append(val)
            out_i.append(torch.where(idx >= 0, idx + self.id_base, idx))
        score, ids = torch.cat(out_s), torch.cat(out_i)
        return self._merge(score, ids, score.shape[1]) if merge else (score, ids)


def _mark(events):
    if events is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


class GraphedSearch:
    """Static-shape, graph-captured hybrid search + router rerank for small query batches (config C2)."""

    def __init__(self, engine: HybridEngine, router, batch: int, max_terms: int, k: int, pool: int):
        if engine.world != 1:
            raise NotImplementedError("graphed_search captures a single-shard step (no collective inside the graph)")
        if not (1 <= batch <= _lib.GEMV_MAX_BATCH):
            raise ValueError(f"graphed_search serves batches of 1..{_lib.GEMV_MAX_BATCH} queries (the GEMV path)")
        if not getattr(router, "stats_initialized", False) and batch > 1:
            per_query = True
        else:
            per_query = not getattr(router, "stats_initialized", False)
        dev = engine.passages.device
        self.engine, self.router, self.batch, self.max_terms, self.k, self.pool = engine, router, batch, max_terms, k, pool
        # static inputs: every query padded to max_terms tokens with -1 (out of vocabulary: contributes nothing)
        self.q_terms = torch.full((batch * max_terms,), -1, dtype=torch.int32, device=dev)
        self.q_off = (torch.arange(batch + 1, dtype=torch.int32, device=dev) * max_terms).contiguous()
        self.q_emb = torch.zeros((batch, engine.passages.shape[1]), dtype=torch.bfloat16, device=dev)
        self._side = torch.cuda.Stream(device=dev)

        def body():
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):                       # branch 2: dense GEMV over the whole shard
                ds, di = ops.dense_gemv_topk(engine.passages, self.q_emb, pool, engine.id_base)
            bs, bi = engine.sparse.score_topk(self.q_terms, self.q_off, max_terms, pool)   # branch 1: seed + BM25 + merge
            cur.wait_stream(self._side)
            ids, sb, sd, sh = ops.hybrid_fuse_topk(bs, bi, ds, di, k)
            fused, order = router.hybrid_rerank(sb, sd, top_k=k, per_query_stats=per_query)
            ranked = torch.gather(ids, 1, order)
            return ranked, fused, torch.gather(sb, 1, order), torch.gather(sd, 1, order)

        with torch.no_grad():
            warm = torch.cuda.Stream(device=dev)                       # warm-up off the default stream, as capture requires
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                for _ in range(2):
                    body()
            torch.cuda.current_stream().wait_stream(warm)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.out = body()
            self.kernels_per_replay = _lib.launch_count() - before    # libragb200 kernels inside the graph

    def load(self, q_terms: Tensor, q_off: Tensor, q_emb: Tensor) -> None:
        """Copy one ragged batch (term ids, offsets, embeddings) into the graph's static, padded buffers."""
        n = q_off.shape[0] - 1
        if n != self.batch:
            raise ValueError(f"this graph was captured for {self.batch} queries, got {n}")
        lens = (q_off[1:] - q_off[:-1]).to(torch.int64)
        if q_terms.numel() == self.batch * self.max_terms and bool((lens == self.max_terms).all()):
            self.q_terms.copy_(q_terms)                                 # already in the padded layout
        else:
            if int(lens.max()) > self.max_terms:
                raise ValueError("a query is longer than the max_terms this graph was captured for")
            self.q_terms.fill_(-1)
            pos = torch.arange(q_terms.shape[0], device=q_terms.device) - torch.repeat_interleave(q_off[:-1].to(torch.int64), lens)
            row = torch.repeat_interleave(torch.arange(n, device=q_terms.device), lens)
            self.q_terms[row * self.max_terms + pos] = q_terms[:int(lens.sum())]
        self.q_emb.copy_(q_emb)

    def replay(self):
        """-> (ids int32 [B,k], fused [B,k], bm25 [B,k], dense [B,k]) of the batch loaded last; the tensors are the
        graph's static outputs (overwritten by the next replay)."""
        self.graph.replay()
        return self.out

    def __call__(self, q_terms: Tensor, q_off: Tensor, q_emb: Tensor):
        self.load(q_terms, q_off, q_emb)
        return self.replay()
