"""CPU restatement of the reference's sparse (BM25) scoring path.  TEST INFRASTRUCTURE ONLY.

Follows
  * rank_bm25 0.2.2 ``BM25Okapi`` (third-party, pinned by ``requirements.txt:8``
    ``rank-bm25>=0.2.2``; not vendored, not installable here).  Call sites in the
    reference: construct ``rag_uq/streaming_index.py:142,220``; score ``:169``.
  * ``BM25Index._tokenize``  ``rag_uq/streaming_index.py:118-120``
  * ``BM25Index.search``     ``rag_uq/streaming_index.py:150-179``

PINNED (``BM25Index.search`` around the scoring call): ``tests/golden/make_retrieval_golden.py``
runs the LIVE ``rag_uq.streaming_index.BM25Index`` from /root/reference (tokeniser, duplicate
skipping, rebuild per add, ``np.argsort(...)[::-1][:top_k]``, the ``> 0`` filter, row -> doc id)
with the module's missing ``BM25Okapi`` bound to ``OkapiLiteral`` and stores its results in
``tests/golden/retrieval_golden.json``; ``index_search`` / ``tokenize`` are checked against it.
PARITY UNPINNED (the rank_bm25 arithmetic itself): the package cannot be installed here and the
reference holds no golden vector for it; the known answers in
``tests/golden/bm25_known_answers.json`` were derived by hand from the published Okapi formula
 and are reproduced by both classes below and by an independent scalar
implementation in ``tests/golden/make_golden.py``.

Two implementations, checked against each other in ``tests/test_oracle_cpu.py``:
  ``OkapiLiteral``  - per-document dict-of-counts, one Python pass over every
                      document per query token: the algorithm as published
                      (and the honest CPU baseline: this is what the reference
                      runs per query).
  ``OkapiCsr``      - the same arithmetic over a term-major CSR built from
                      integer term ids with numpy; float64 like the original.
                      Used where the literal form would take minutes.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

K1_DEFAULT = 1.5      # streaming_index.py:103
B_DEFAULT = 0.75      # streaming_index.py:104
EPSILON_DEFAULT = 0.25  # rank_bm25 BM25Okapi default, never overridden by the reference


def tokenize(text: str) -> List[str]:
    """streaming_index.py:118-120 - lower-case, split on whitespace, duplicates kept."""
    return text.lower().split()


class OkapiLiteral:
    """Dict-based Okapi BM25 exactly as rank_bm25 0.2.2 computes it."""

    def __init__(self, corpus: Sequence[Sequence], k1: float = K1_DEFAULT, b: float = B_DEFAULT,
                 epsilon: float = EPSILON_DEFAULT):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = len(corpus)
        self.doc_len = [len(doc) for doc in corpus]
        self.doc_freqs: List[Dict] = []
        containing: Dict = {}            # term -> number of documents holding it
        for doc in corpus:
            counts: Dict = {}
            for tok in doc:
                counts[tok] = counts.get(tok, 0) + 1
            self.doc_freqs.append(counts)
            for tok in counts:
                containing[tok] = containing.get(tok, 0) + 1
        self.avgdl = sum(self.doc_len) / self.corpus_size
        # idf = ln(N - n + 0.5) - ln(n + 0.5); every NEGATIVE idf is replaced by
        # epsilon * mean(idf), the mean taken over the raw values (negatives included).
        self.idf: Dict = {}
        total = 0.0
        below_zero = []
        for tok, n in containing.items():
            val = math.log(self.corpus_size - n + 0.5) - math.log(n + 0.5)
            self.idf[tok] = val
            total += val
            if val < 0:
                below_zero.append(tok)
        self.average_idf = total / len(self.idf)
        floor = self.epsilon * self.average_idf
        for tok in below_zero:
            self.idf[tok] = floor

    def get_scores(self, query_tokens: Sequence) -> np.ndarray:
        """One float64 score per document; every token OCCURRENCE contributes."""
        out = np.zeros(self.corpus_size)
        dl = np.array(self.doc_len)
        for tok in query_tokens:
            tf = np.array([(d.get(tok) or 0) for d in self.doc_freqs])
            weight = self.idf.get(tok) or 0
            out += weight * (tf * (self.k1 + 1) / (tf + self.k1 * (1 - self.b + self.b * dl / self.avgdl)))
        return out


def index_search(scores: np.ndarray, top_k: int) -> List[Tuple[int, float]]:
    """``BM25Index.search`` after scoring (streaming_index.py:171-179).

    The reference takes ``np.argsort(scores)[::-1][:top_k]`` (tie order is an
    artefact of introsort) and keeps entries with score > 0.  We fix the tie
    order the framework promises instead: score descending, then index ascending.
    Returns (doc index, float score) pairs.
    """
    n = scores.shape[0]
    order = np.lexsort((np.arange(n), -scores))[:top_k]
    return [(int(i), float(scores[i])) for i in order if scores[i] > 0]


class OkapiCsr:
    """Same arithmetic over integer term ids with a term-major CSR (float64)."""

    def __init__(self, doc_off: np.ndarray, doc_tok: np.ndarray, vocab: int,
                 k1: float = K1_DEFAULT, b: float = B_DEFAULT, epsilon: float = EPSILON_DEFAULT):
        doc_off = np.asarray(doc_off, dtype=np.int64)
        doc_tok = np.asarray(doc_tok, dtype=np.int64)
        self.k1, self.b, self.epsilon = k1, b, epsilon
        n = doc_off.shape[0] - 1
        self.corpus_size = n
        self.vocab = vocab
        self.doc_len = np.diff(doc_off)
        self.avgdl = float(self.doc_len.sum()) / n
        owner = np.repeat(np.arange(n, dtype=np.int64), self.doc_len)
        pair, tf = np.unique(doc_tok * n + owner, return_counts=True)
        self.post_doc = (pair % n).astype(np.int32)
        self.post_tf = tf.astype(np.int32)
        df = np.bincount(pair // n, minlength=vocab).astype(np.int64)
        self.df = df
        self.term_off = np.concatenate([[0], np.cumsum(df)]).astype(np.int64)
        self.idf, self.average_idf = okapi_idf(df, n, epsilon)

    def get_scores(self, query_terms: Sequence[int]) -> np.ndarray:
        out = np.zeros(self.corpus_size)
        for t in query_terms:
            if t < 0 or t >= self.vocab:
                continue                      # OOV: idf.get(q) is None -> contributes 0
            lo, hi = self.term_off[t], self.term_off[t + 1]
            docs = self.post_doc[lo:hi]
            tf = self.post_tf[lo:hi].astype(np.float64)
            dl = self.doc_len[docs]
            out[docs] += self.idf[t] * (tf * (self.k1 + 1) / (tf + self.k1 * (1 - self.b + self.b * dl / self.avgdl)))
        return out


def okapi_idf(df: np.ndarray, corpus_size: int, epsilon: float = EPSILON_DEFAULT):
    """idf[V] (float64) with the epsilon floor; terms with df == 0 are absent -> 0."""
    df = np.asarray(df, dtype=np.int64)
    present = df > 0
    raw = np.zeros(df.shape[0])
    raw[present] = np.log(corpus_size - df[present] + 0.5) - np.log(df[present] + 0.5)
    average = float(raw[present].sum()) / max(int(present.sum()), 1)
    idf = raw.copy()
    idf[present & (raw < 0)] = epsilon * average
    return idf, average
