"""Exact check of a handful of queries against a corpus too large for the dict / CSR oracles.  TEST INFRASTRUCTURE ONLY.

Used by ``bench.py --verify`` (outside the timed region) and by the slow ``-m gpu`` tests: the north-star asks
for "exact-match top-10 ids vs. the reference over a 10M x 768 synthetic corpus", and ``OkapiLiteral`` /
``OkapiCsr`` cannot hold 1.5 G tokens.  The arithmetic is the same restatement, streamed:

* BM25       rank_bm25 0.2.2 ``BM25Okapi.get_scores`` as called from ``BM25Index.search``
             (rag_uq/streaming_index.py:169): float64, one term of the sum per query-token OCCURRENCE, idf with
             the epsilon floor from ``oracle.bm25_okapi.okapi_idf``; then ``index_search`` (:171-179).
* dense      exact inner products of the bf16-rounded unit rows in float64 (stand-in for :355-370).
* fusion     ``oracle.dense_fusion.hybrid_search`` / ``scores_for_router`` (:464-557).
* rerank     ``oracle.router.hybrid_rerank`` (router.py:179-202; call site run_evaluation.py:165-184).

The corpus is regenerated from its seed chunk by chunk.  Integer work that only SELECTS data (which tokens of a
chunk are query terms, distinct (document, term) pairs for the document frequencies) may run through torch on
any device - the synthetic generator is integer-exact on CPU and CUDA (tests/test_host_cpu.py) - every floating
point operation of the BM25 / fusion check happens here in numpy float64 on the host.  Dense inner products are a
plain torch float64 matmul of the chunk (the "plain torch reference" for a floating-point kernel).
Nothing here imports the product package: the generator module is handed in by the caller.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import bm25_okapi, dense_fusion


def top_desc(scores: np.ndarray, k: int, positive_only: bool = False):
    """Exact top-k of a long vector in the framework's order (score desc, index asc) without sorting all of it.
    Returns (indices, scores, gap) where gap = relative distance between the last kept and the first dropped
    score (inf when nothing was dropped): a tiny gap marks a cut that fp32 and fp64 may place differently."""
    n = scores.shape[0]
    k_eff = min(k, n)
    if k_eff == 0:
        return np.zeros(0, np.int64), np.zeros(0), np.inf
    take = min(n, k_eff + 64)
    cand = np.argpartition(-scores, take - 1)[:take] if take < n else np.arange(n)
    if take < n:                                   # everything tied with the partition boundary must be considered
        edge = scores[cand].min()
        extra = np.nonzero(scores == edge)[0]
        cand = np.union1d(cand, extra)
    order = cand[np.lexsort((cand, -scores[cand]))]
    kept = order[:k_eff]
    gap = np.inf
    if order.shape[0] > k_eff:
        a, b = scores[kept[-1]], scores[order[k_eff]]
        gap = abs(a - b) / max(abs(a), 1e-300)
    if positive_only:
        kept = kept[scores[kept] > 0]
    return kept, scores[kept], gap


class LargeHybridCheck:
    """Streams chunks of the corpus and keeps full float64 score vectors for a few queries."""

    def __init__(self, df: np.ndarray, n_docs: int, total_len: int, query_terms: Sequence[Sequence[int]],
                 k1: float = bm25_okapi.K1_DEFAULT, b: float = bm25_okapi.B_DEFAULT,
                 epsilon: float = bm25_okapi.EPSILON_DEFAULT):
        self.k1, self.b = k1, b
        self.n_docs = int(n_docs)
        self.avgdl = float(total_len) / self.n_docs                    # rank_bm25: sum(len) / corpus_size
        self.vocab = int(df.shape[0])
        self.idf, self.average_idf = bm25_okapi.okapi_idf(df, self.n_docs, epsilon)
        self.query_terms = [[int(t) for t in q] for q in query_terms]
        distinct = sorted({t for q in self.query_terms for t in q if 0 <= t < self.vocab})
        self.terms = np.asarray(distinct, dtype=np.int64)              # the caller selects occurrences of these
        self.slot: Dict[int, int] = {t: i for i, t in enumerate(distinct)}
        n_q = len(self.query_terms)
        self.bm25 = np.zeros((n_q, self.n_docs))
        self.dense = np.zeros((n_q, self.n_docs))
        self.rows_seen = 0

    def add_chunk(self, row0: int, doc_len: np.ndarray, occ_doc: np.ndarray, occ_slot: np.ndarray,
                  dense_chunk: np.ndarray) -> None:
        """``occ_doc`` / ``occ_slot``: one entry per TOKEN of the chunk that is one of ``self.terms`` (document
        index inside the chunk, index into ``self.terms``); ``dense_chunk`` float64 [n_queries, n]."""
        n = int(doc_len.shape[0])
        n_t = max(len(self.terms), 1)
        tf = np.bincount(occ_doc.astype(np.int64) * n_t + occ_slot.astype(np.int64), minlength=n * n_t).reshape(n, n_t)
        dl = doc_len.astype(np.float64)
        denom_len = self.k1 * (1 - self.b + self.b * dl / self.avgdl)
        for q, toks in enumerate(self.query_terms):
            out = self.bm25[q, row0:row0 + n]
            for t in toks:                                            # every OCCURRENCE adds (duplicates count twice)
                if t not in self.slot:
                    continue                                          # OOV: idf.get(q) is None -> 0
                f = tf[:, self.slot[t]].astype(np.float64)
                out += self.idf[t] * (f * (self.k1 + 1) / (f + denom_len))
        self.dense[:, row0:row0 + n] = dense_chunk
        self.rows_seen += n

    def finish(self, pool: int, k: int, router_state=None, stats_initialized: bool = True, candidates: int = 0) -> List[dict]:
        """Per query: the pools, the fused top-``max(k, candidates)`` and the router-reranked top-k, plus the
        relative gaps at the two pool cuts (see ``top_desc``)."""
        assert self.rows_seen == self.n_docs, (self.rows_seen, self.n_docs)
        out = []
        for q in range(len(self.query_terms)):
            bi, bs, bgap = top_desc(self.bm25[q], pool, positive_only=True)
            di, ds, dgap = top_desc(self.dense[q], pool)
            bm = [(int(i), float(s)) for i, s in zip(bi, bs)]
            de = [(int(i), float(s)) for i, s in zip(di, ds)]
            n_c = max(k, candidates)
            fused = dense_fusion.hybrid_search(bm, de, n_c)
            rec = {"bm25_pool": bm, "dense_pool": de, "bm25_gap": float(bgap), "dense_gap": float(dgap), "fused": fused}
            if router_state is not None:
                import torch

                from . import router as router_oracle
                sb, sd, ids = dense_fusion.scores_for_router(bm, de, n_c)
                with torch.no_grad():
                    vals, order = router_oracle.hybrid_rerank(torch.tensor([sb], dtype=torch.float32),
                                                              torch.tensor([sd], dtype=torch.float32), router_state,
                                                              stats_initialized, k)
                    w = router_oracle.gate(torch.tensor([sb], dtype=torch.float32), torch.tensor([sd], dtype=torch.float32),
                                           router_state, stats_initialized)
                    every = router_oracle.fused(torch.tensor([sb], dtype=torch.float32),
                                                torch.tensor([sd], dtype=torch.float32), w)[0].tolist()
                rec["rerank_ids"] = [ids[j] for j in order[0].tolist()]
                rec["rerank_vals"] = vals[0].tolist()
                rec["rerank_all"] = {i: v for i, v in zip(ids, every) if i >= 0}     # fused score of every candidate
            out.append(rec)
        return out


def compare_ranking(got_ids: Sequence[int], got_vals: Sequence[float], want_ids: Sequence[int], want_vals: Sequence[float],
                    all_vals: Dict[int, float], rtol: float, atol: float):
    """-> (exact, explained): ``exact`` = identical id lists; ``explained`` = every deviation is a (near-)tie in
    the ORACLE's own scores (the returned id's oracle score is within tolerance of the oracle score at that rank)."""
    got_ids = [int(i) for i in got_ids if int(i) >= 0]
    want_ids = [int(i) for i in want_ids if int(i) >= 0]
    if got_ids == want_ids:
        ok = bool(np.allclose(got_vals[:len(want_ids)], want_vals[:len(want_ids)], rtol=rtol, atol=atol))
        return ok, ok
    if len(got_ids) != len(want_ids) or len(set(got_ids)) != len(got_ids):
        return False, False
    for j, gi in enumerate(got_ids):
        truth = all_vals.get(gi)
        if truth is None or abs(truth - want_vals[j]) > rtol * abs(want_vals[j]) + atol:
            return False, False
        if abs(got_vals[j] - want_vals[j]) > rtol * abs(want_vals[j]) + atol:
            return False, False
    return False, True


def run_synthetic_check(synth, device, n_passages: int, dim: int, q_terms_rows: Sequence[Sequence[int]], q_emb,
                        pool: int, k: int, router_state=None, stats_initialized: bool = True, candidates: int = 0,
                        chunk_docs: int = 250_000, df_expect=None):
    """Regenerate the synthetic corpus of ``synth`` (the product's generator module, handed in) chunk by chunk and
    run ``LargeHybridCheck`` for the given queries.  ``q_emb``: torch bf16 [n_q, dim] on ``device``.
    ``df_expect`` (optional int tensor [V]): the document frequencies the product scored with; they must be equal
    to the ones counted here.  Returns (per-query records, info dict)."""
    import torch

    vocab = synth.vocab_size(n_passages)
    cdf = synth.zipf_cdf(vocab, device)
    # pass 1: document frequencies, document count, total length (integer counting only)
    df = torch.zeros(vocab, dtype=torch.int64, device=device)
    total_len = 0
    for b0 in range(0, n_passages, chunk_docs):
        b1 = min(n_passages, b0 + chunk_docs)
        doc_off, tok = synth.doc_tokens(b0, b1, cdf)
        lens = doc_off[1:] - doc_off[:-1]
        owner = torch.repeat_interleave(torch.arange(b1 - b0, device=device, dtype=torch.int64), lens)
        pairs = torch.unique(owner * vocab + tok.to(torch.int64))
        df += torch.bincount(pairs % vocab, minlength=vocab)
        total_len += int(lens.sum())
        del doc_off, tok, lens, owner, pairs
    info = {"df_matches_product": None}
    if df_expect is not None:
        info["df_matches_product"] = bool(torch.equal(df, df_expect.to(device=device, dtype=torch.int64)))
    check = LargeHybridCheck(df.cpu().numpy(), n_passages, total_len, q_terms_rows)
    terms_dev = torch.as_tensor(check.terms, device=device)
    q64 = q_emb.to(device=device, dtype=torch.float64)
    # pass 2: term occurrences of the query terms + float64 inner products, chunk by chunk
    for b0 in range(0, n_passages, chunk_docs):
        b1 = min(n_passages, b0 + chunk_docs)
        doc_off, tok = synth.doc_tokens(b0, b1, cdf)
        lens = doc_off[1:] - doc_off[:-1]
        owner = torch.repeat_interleave(torch.arange(b1 - b0, device=device, dtype=torch.int32), lens)
        if terms_dev.numel():
            slot = torch.searchsorted(terms_dev, tok.to(torch.int64)).clamp_(max=terms_dev.numel() - 1)
            hit = terms_dev[slot] == tok.to(torch.int64)
            occ_doc, occ_slot = owner[hit].cpu().numpy(), slot[hit].to(torch.int32).cpu().numpy()
        else:
            occ_doc, occ_slot = np.zeros(0, np.int32), np.zeros(0, np.int32)
        rows = synth.passage_embeddings(b0, b1, dim, device)
        dense = (q64 @ rows.to(torch.float64).T).cpu().numpy()
        check.add_chunk(b0, lens.cpu().numpy(), occ_doc, occ_slot, dense)
        del doc_off, tok, lens, owner, rows
    info.update({"avgdl": check.avgdl, "average_idf": check.average_idf, "distinct_query_terms": int(len(check.terms))})
    return check.finish(pool, k, router_state, stats_initialized, candidates), info
