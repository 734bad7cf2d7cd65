"""Seeded synthetic corpora and queries.

Everything is a pure function of (seed, global row, position) through a 64-bit integer hash
evaluated with torch integer ops, so any shard of the corpus can be generated on any device
without generating the rest, and a query can re-derive "its" passage without the corpus being
resident.  No transcendental functions are used, so CPU and CUDA produce identical integers.

  passages    iid approximately-normal vectors (sum of four 16-bit uniforms), unit-normalised
              in fp32, rounded to bf16 - the bf16 tensor is the ground truth for oracle and
              kernel alike.
  tokens      document length clip(round(N(150, 40)), 20, 300) (the reference chunker emits
              200-word passages with 50 overlap, data/preprocessing/prepare_corpus.py:28-34);
              term ids drawn from Zipf(s = 1) over V = min(5M, max(50k, N/2)) terms.
  queries     8 tokens sampled with replacement from one random passage (so matches exist and
              duplicates happen), 1% of the queries get one out-of-vocabulary id; the query
              embedding is that passage's vector plus N(0, 0.5^2)/sqrt(dim) noise, re-normalised.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import torch
from torch import Tensor

CORPUS_SEED = 1234
QUERY_SEED = 4321
_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C0, _C1, _C2 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def _lsr(x: Tensor, s: int) -> Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def mix64(x: Tensor) -> Tensor:
    """splitmix64 finaliser on int64 tensors (two's-complement wrap-around arithmetic)."""
    x = x + _C0
    x = (x ^ _lsr(x, 30)) * _C1
    x = (x ^ _lsr(x, 27)) * _C2
    return x ^ _lsr(x, 31)


def hash3(seed: int, a: Tensor, b: Tensor | int) -> Tensor:
    h = mix64(a.to(torch.int64) + _s64(seed * 0x632BE59BD9B4E019))
    return mix64(h ^ (torch.as_tensor(b, device=a.device).to(torch.int64) * _s64(0xD6E8FEB86659FD93)))


def uniform01(h: Tensor) -> Tensor:
    """float64 in [0, 1) from the top 53 bits."""
    return _lsr(h, 11).to(torch.float64) * (2.0 ** -53)


def approx_normal(h: Tensor) -> Tensor:
    """Irwin-Hall(4) from the four 16-bit fields of one hash: mean 0, variance 1 (float32, exact integers inside)."""
    s = (h & 0xFFFF) + (_lsr(h, 16) & 0xFFFF) + (_lsr(h, 32) & 0xFFFF) + (_lsr(h, 48) & 0xFFFF)
    return (s.to(torch.float32) - 131070.0) * (1.0 / 37837.0)  # sqrt(4 * (65536^2 - 1) / 12) = 37837.2


def vocab_size(n_passages: int) -> int:
    return min(5_000_000, max(50_000, n_passages // 2))


def zipf_cdf(vocab: int, device) -> Tensor:
    w = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64)
    cdf = torch.cumsum(w, 0)
    return (cdf / cdf[-1]).to(device)


def passage_embeddings(row_begin: int, row_end: int, dim: int, device, seed: int = CORPUS_SEED,
                       chunk_rows: int = 1 << 17) -> Tensor:
    """bf16 [row_end - row_begin, dim], unit rows."""
    out = torch.empty((row_end - row_begin, dim), dtype=torch.bfloat16, device=device)
    cols = torch.arange(dim, device=device, dtype=torch.int64)
    for r0 in range(row_begin, row_end, chunk_rows):
        r1 = min(row_end, r0 + chunk_rows)
        rows = torch.arange(r0, r1, device=device, dtype=torch.int64)
        x = approx_normal(hash3(seed, rows[:, None] * 4096 + cols[None, :], 0x51))
        x = x / x.norm(dim=1, keepdim=True)
        out[r0 - row_begin:r1 - row_begin] = x.to(torch.bfloat16)
    return out


def doc_lengths(row_begin: int, row_end: int, device, seed: int = CORPUS_SEED) -> Tensor:
    rows = torch.arange(row_begin, row_end, device=device, dtype=torch.int64)
    g = approx_normal(hash3(seed, rows, 0xD0C))
    return torch.clamp(torch.round(150.0 + 40.0 * g), 20, 300).to(torch.int64)


def doc_tokens(row_begin: int, row_end: int, cdf: Tensor, seed: int = CORPUS_SEED) -> Tuple[Tensor, Tensor]:
    """(doc_off int64 [n+1], doc_tok int32 [total]) for global rows [row_begin, row_end)."""
    device = cdf.device
    lens = doc_lengths(row_begin, row_end, device, seed)
    doc_off = torch.zeros(lens.shape[0] + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=doc_off[1:])
    total = int(doc_off[-1])
    owner = torch.repeat_interleave(torch.arange(row_begin, row_end, device=device, dtype=torch.int64), lens)
    pos = torch.arange(total, device=device, dtype=torch.int64) - doc_off[:-1].repeat_interleave(lens)
    u = uniform01(hash3(seed, owner * 512 + pos, 0x70C))
    tok = torch.searchsorted(cdf, u, right=True).clamp_(max=cdf.shape[0] - 1).to(torch.int32)
    return doc_off, tok


@dataclass
class QueryBatch:
    q_terms: Tensor      # int32 [total]
    q_off: Tensor        # int32 [B+1]
    q_emb: Tensor        # bf16 [B, dim]
    source_rows: Tensor  # int64 [B] the passage each query was derived from
    max_terms: int


def make_queries(n_queries: int, n_passages: int, dim: int, cdf: Tensor, device, first_query: int = 0,
                 n_terms: int = 8, seed: int = QUERY_SEED, corpus_seed: int = CORPUS_SEED) -> QueryBatch:
    qi = torch.arange(first_query, first_query + n_queries, device=device, dtype=torch.int64)
    src = (_lsr(hash3(seed, qi, 0x5C), 1) % n_passages)
    # tokens: n_terms positions sampled with replacement from the source passage's token list
    g = approx_normal(hash3(corpus_seed, src, 0xD0C))
    lens = torch.clamp(torch.round(150.0 + 40.0 * g), 20, 300).to(torch.int64)
    slots = torch.arange(n_terms, device=device, dtype=torch.int64)
    pos = _lsr(hash3(seed, qi[:, None] * 64 + slots[None, :], 0x7E), 1) % lens[:, None]
    u = uniform01(hash3(corpus_seed, src[:, None] * 512 + pos, 0x70C))
    terms = torch.searchsorted(cdf, u, right=True).clamp_(max=cdf.shape[0] - 1).to(torch.int32)
    oov = (_lsr(hash3(seed, qi, 0x00F), 1) % 100) == 0
    terms[:, n_terms - 1] = torch.where(oov, torch.full_like(terms[:, 0], cdf.shape[0] + 7), terms[:, n_terms - 1])
    q_off = (torch.arange(n_queries + 1, device=device, dtype=torch.int64) * n_terms).to(torch.int32)
    # embedding: source passage vector + noise
    cols = torch.arange(dim, device=device, dtype=torch.int64)
    base = approx_normal(hash3(corpus_seed, src[:, None] * 4096 + cols[None, :], 0x51))
    base = base / base.norm(dim=1, keepdim=True)
    noise = approx_normal(hash3(seed, qi[:, None] * 4096 + cols[None, :], 0xE0)) * (0.5 / math.sqrt(dim))
    q = base + noise
    q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    return QueryBatch(terms.reshape(-1).contiguous(), q_off, q.contiguous(), src, n_terms)


def build_synthetic_engine(n_passages: int, dim: int, device, rank: int = 0, world: int = 1, group=None,
                           block_docs: int = 250_000, mma_variant: int = 3, with_sparse: bool = True,
                           with_dense: bool = True):
    """Generate this rank's row shard of the synthetic corpus and wrap it in a HybridEngine.

    Returns (engine, cdf).  BM25 statistics are global: df, N and the total length are
    all-reduced over ``group`` before idf / norm are computed.
    """
    from .engine import HybridEngine, global_bm25_statistics, shard_rows
    from .sparse import build_shard_blocked

    lo, hi = shard_rows(n_passages, world, rank)
    vocab = vocab_size(n_passages)
    cdf = zipf_cdf(vocab, device)
    passages = passage_embeddings(lo, hi, dim, device) if with_dense else None
    sparse = None
    if with_sparse:
        def blocks():
            for b0 in range(lo, hi, block_docs):
                yield doc_tokens(b0, min(hi, b0 + block_docs), cdf)
        sparse = build_shard_blocked(blocks(), hi - lo, vocab, device, id_base=lo)
        df, n_all, len_all = global_bm25_statistics(sparse.df, hi - lo, int(sparse.doc_len.sum()), group)
        sparse.finalize(df, n_all, len_all, group)
    return HybridEngine(sparse, passages, id_base=lo, group=group, mma_variant=mma_variant), cdf
