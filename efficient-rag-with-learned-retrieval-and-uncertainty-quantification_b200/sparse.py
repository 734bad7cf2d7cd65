"""Term-major CSR inverted index resident in HBM, and its builders.

GPU equivalent of what ``BM25Okapi.__init__`` builds per ``BM25Index.add_documents`` /
``_load`` (rag_uq/streaming_index.py:140-142, 219-220): per-document term counts, document
lengths, average length and idf with the epsilon floor.  Raw term frequencies and lengths are
kept (not baked impacts), so global statistics can be refreshed in O(V + N) when documents
are added or when shards exchange their document frequencies.

Layout (one shard = a contiguous range of global passage rows):
    term_off [V+1] int64   postings of term t are [term_off[t], term_off[t+1])
    post_doc [nnz] int32   LOCAL row of the posting, ascending inside a term
    post_tf  [nnz] int16   bit pattern of a uint16 term frequency (clipped at 65535)
    doc_len  [N]   int32
    df       [V]   int32   LOCAL document frequency (summed over shards -> global)
    idf      [V]   float32, norm [N] float32   set by ``finalize``
The sorting / segmenting below is torch plumbing (sort, unique_consecutive, cumsum); the
statistics and all scoring run in the library's own kernels.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import ops

MAX_DENSE_TERMS = 1024          # rows the kernel's table directory can hold
DENSE_MIN_FRACTION = int(os.environ.get("RAGB_DENSE_MIN_FRACTION", "24"))  # a term gets a row when df >= N / this ...
DENSE_TABLE_BYTES = 8 << 30     # ... while the table stays under this many bytes per shard
IMPACT_TABLE_BYTES = 12 << 30   # the fp16 impact bounds of the table rows are optional: skipped above this size
POSTING_IMPACTS = int(os.environ.get("RAGB_POSTING_IMPACTS", "1"))   # bake tf / (tf + norm) per posting (4 bytes each) for the window phase of the search
IMPACT_CAP_TAIL = float(os.environ.get("RAGB_IMPACT_CAP_TAIL", "0.0001"))   # share of a row's documents listed above its cap (0 = no cap)


@dataclass
class SparseShard:
    term_off: Tensor
    post_doc: Tensor
    post_tf: Tensor
    doc_len: Tensor
    df: Tensor
    n_docs: int
    vocab: int
    id_base: int = 0
    k1: float = 1.5
    b: float = 0.75
    epsilon: float = 0.25
    idf: Optional[Tensor] = None
    norm: Optional[Tensor] = None
    corpus_size: int = 0
    avgdl: float = 0.0
    dense_tf: Optional[Tensor] = None      # uint8 [n_dense, stride]: tf rows of the most frequent terms
    dense_terms: Optional[Tensor] = None   # int32 [n_dense]
    dense_imp: Optional[Tensor] = None     # float16 [n_dense, stride]: upper bounds of tf / (tf + norm)
    dense_maximp: Optional[Tensor] = None  # float32 [n_dense]: their row maxima
    dense_cap: Optional[Tensor] = None     # float32 [n_dense]: impact cap of each row (all but the marker documents stay below)
    hi_off: Optional[Tensor] = None        # int32 [n_dense + 1]
    hi_doc: Optional[Tensor] = None        # int32: ascending local rows of the documents above the cap, row by row
    post_imp: Optional[Tensor] = None      # float32 [nnz]: tf / (tf + norm[doc]) of every posting (derived; rebuilt by finalize)
    use_dense_table: bool = True
    bake_impacts: bool = True              # SegmentedIndex turns it off: its refresh stays O(V + N), not O(nnz)
    df_global: Optional[Tensor] = None     # document frequencies summed over all shards (set by finalize)

    @property
    def nnz(self) -> int:
        return int(self.post_doc.shape[0])

    def finalize(self, df_global: Optional[Tensor] = None, corpus_size: Optional[int] = None,
                 total_len: Optional[int] = None, group=None) -> "SparseShard":
        """(Re)compute idf[V] and norm[N] from GLOBAL statistics (defaults: this shard alone)."""
        df_global = self.df if df_global is None else df_global
        self.df_global = df_global           # what idf was computed from (bench --verify compares it with its own count)
        self.corpus_size = self.n_docs if corpus_size is None else int(corpus_size)
        total = int(self.doc_len.sum()) if total_len is None else int(total_len)
        self.avgdl = total / self.corpus_size
        self.idf = ops.bm25_build_idf(df_global.to(torch.int32), self.corpus_size, self.epsilon)
        self.norm = ops.bm25_build_norm(self.doc_len, self.avgdl, self.k1, self.b)
        self._build_dense_table(df_global, group)
        self._build_impact_bounds()
        # baked impacts: norm moves with the global statistics, so they are rebuilt here (one streaming kernel)
        self.post_imp = None
        if POSTING_IMPACTS and self.bake_impacts and self.nnz > 0 and self.post_doc.is_cuda:
            self.post_imp = ops.bm25_build_posting_impacts(self.post_doc, self.post_tf, self.norm)
        return self

    def _build_impact_bounds(self) -> None:
        """fp16 upper bounds of tf / (tf + norm) for the table terms and their row maxima (ragb200.h): they let
        the kernel bound what the table terms can add to a document and mark, in a cheap fp16 pass, the few
        documents of a super-range that are worth the exact arithmetic.  norm changes with the global
        statistics, so this runs after every finalize."""
        dev = self.post_doc.device
        self.dense_imp = torch.empty(0, dtype=torch.float16, device=dev)
        self.dense_maximp = torch.empty(0, dtype=torch.float32, device=dev)
        if self.dense_tf is None or self.dense_tf.numel() == 0 or self.dense_tf.numel() * 2 > IMPACT_TABLE_BYTES:
            return
        self.dense_imp, self.dense_maximp = ops.bm25_build_impact_bounds(self.dense_tf, self.norm)
        self._build_impact_cap()

    def _build_impact_cap(self) -> None:
        """Impact cap of the table rows (ragb200.h): the row maximum that bounds what a table term can add to a document
        is set by a handful of documents (tf 15 of a stop-word in a short passage).  Capping every row at its
        (1 - IMPACT_CAP_TAIL) quantile and listing the few documents above the cap as mark-only "marker lists" tightens
        the bound for everybody else: at 10M passages the share of queries whose threshold ends above the table bound -
        the condition for the cheap window phase - rises from 93.5 % to 98.9 % for +9 % postings
        (scripts/analyze_bm25_bounds.py).  torch plumbing (one top-k per row); the bound is per shard, results do not
        depend on it."""
        dev = self.post_doc.device
        self.dense_cap = torch.empty(0, dtype=torch.float32, device=dev)
        self.hi_off = torch.empty(0, dtype=torch.int32, device=dev)
        self.hi_doc = torch.empty(0, dtype=torch.int32, device=dev)
        if IMPACT_CAP_TAIL <= 0 or self.dense_imp is None or self.dense_imp.numel() == 0 or self.n_docs < 4096:
            return
        rows = self.dense_imp.shape[0]
        keep = max(1, int(round(IMPACT_CAP_TAIL * self.n_docs)))
        caps, lists, offs = [], [], [0]
        for r in range(rows):
            top = torch.topk(self.dense_imp[r, :self.n_docs], keep + 1)
            cap = top.values[keep]                                    # the (keep + 1)-th largest bound of the row
            above = top.indices[:keep][top.values[:keep] > cap]       # strictly above: everything else is <= cap
            lists.append(torch.sort(above).values.to(torch.int32))
            offs.append(offs[-1] + int(above.numel()))
            caps.append(cap.to(torch.float32))
        self.dense_cap = torch.stack(caps).contiguous()
        self.hi_off = torch.tensor(offs, dtype=torch.int32, device=dev)
        self.hi_doc = torch.cat(lists).contiguous() if offs[-1] else torch.zeros(1, dtype=torch.int32, device=dev)

    def _build_dense_table(self, df_global: Tensor, group=None) -> None:
        """Dense uint8 tf rows for the terms present in >= 1/64 of ALL documents (capped by memory).

        The choice uses global document frequencies and a cross-shard agreement on "every tf fits
        a byte", so all shards pick the same terms and accumulate in the same order.
        """
        dev = self.post_doc.device
        keep = (self.dense_tf, self.dense_terms) if (self.dense_terms is not None and self.dense_terms.numel()) else None
        empty_u8 = torch.empty(0, dtype=torch.uint8, device=dev)
        self.dense_tf, self.dense_terms = empty_u8, torch.empty(0, dtype=torch.int32, device=dev)
        if not self.use_dense_table or self.n_docs == 0:
            return
        cand = torch.nonzero(df_global.to(torch.int64) * DENSE_MIN_FRACTION >= self.corpus_size).flatten()
        if cand.numel() == 0:
            return
        if keep is not None and keep[1].numel() == cand.numel() and torch.equal(keep[1].to(torch.int64), cand):
            self.dense_tf, self.dense_terms = keep     # same terms as before: the rows are still right
            return
        stride = (self.n_docs + 255) // 256 * 256
        # the row budget depends only on global quantities, so every shard keeps the same terms
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rows_max = max(1, min(MAX_DENSE_TERMS, DENSE_TABLE_BYTES // max(256, -(-self.corpus_size // world))))
        if cand.numel() > rows_max:
            top = torch.topk(df_global[cand].to(torch.int64), rows_max).indices
            cand = cand[top].sort().values
        # every tf of a table term must fit a byte on EVERY shard (one kernel + one MIN all-reduce, no host loop)
        fits = (ops.bm25_term_max_tf(self.term_off, self.post_tf, cand.to(torch.int32)) <= 255).to(torch.int32)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(fits, op=dist.ReduceOp.MIN, group=group)
        cand = cand[fits.bool()]
        if cand.numel() == 0:
            return
        self.dense_terms = cand.to(torch.int32).contiguous()
        self.dense_tf = ops.bm25_build_dense_table(self.term_off, self.post_doc, self.post_tf, self.dense_terms, self.n_docs, stride)

    def score_topk(self, q_terms: Tensor, q_off: Tensor, max_terms: int, k: int, seed: Optional[Tensor] = None):
        """``seed`` (float32 [B], optional): proven lower bounds of every query's k-th best score, e.g. ``self.seed``
        raised to the maximum over all shards; None = the kernel seeds itself."""
        if seed is None:
            seed = torch.empty(0, dtype=torch.float32, device=self.post_doc.device)
        cap, hoff, hdoc = self._cap_tensors()
        return ops.bm25_score_topk(self.term_off, self.post_doc, self.post_tf, self.norm, self.idf, self.k1,
                                   self.dense_tf, self.dense_terms, self.dense_imp, self.dense_maximp, q_terms, q_off,
                                   max_terms, self.id_base, k, seed, cap, hoff, hdoc, self._impacts())

    def _impacts(self) -> Tensor:
        if self.post_imp is None:
            return torch.empty(0, dtype=torch.float32, device=self.post_doc.device)
        return self.post_imp

    def _cap_tensors(self):
        if self.dense_cap is None or self.hi_off is None or self.hi_doc is None:
            e = torch.empty(0, dtype=torch.float32, device=self.post_doc.device)
            return e, e.to(torch.int32), e.to(torch.int32)
        return self.dense_cap, self.hi_off, self.hi_doc

    def score_part(self, q_terms: Tensor, q_off: Tensor, max_terms: int, k: int, workspace: Tensor, stripe_begin: int,
                   stripe_end: int, min_smem_bytes: int = 0, seed: Optional[Tensor] = None) -> None:
        """Stripes [stripe_begin, stripe_end) of the staged search (``ops.bm25_stripe_count`` stripes in all; the part
        starting at 0 first; ``ops.bm25_score_finish`` merges)."""
        if seed is None:
            seed = torch.empty(0, dtype=torch.float32, device=self.post_doc.device)
        cap, hoff, hdoc = self._cap_tensors()
        ops.bm25_score_part(self.term_off, self.post_doc, self.post_tf, self.norm, self.idf, self.k1, self.dense_tf,
                            self.dense_terms, self.dense_imp, self.dense_maximp, q_terms, q_off, max_terms, self.id_base, k,
                            seed, cap, hoff, hdoc, self._impacts(), stripe_begin, stripe_end, min_smem_bytes, workspace)

    def score_docs(self, q_terms: Tensor, q_off: Tensor, max_terms: int, cand_ids: Tensor) -> Tensor:
        """Exact BM25 scores [B, C] of chosen passages (global ids, -1 = none): ragb_bm25_score_docs."""
        return ops.bm25_score_docs(self.term_off, self.post_doc, self.post_tf, self.norm, self.idf, self.k1, self.dense_tf,
                                   self.dense_terms, q_terms, q_off, max_terms, self.id_base, cand_ids)

    def seed(self, q_terms: Tensor, q_off: Tensor, max_terms: int, k: int) -> Tensor:
        """Proven lower bounds [B] of the k-th best score of every query over this shard (ragb_bm25_seed)."""
        return ops.bm25_seed(self.term_off, self.post_doc, self.post_tf, self.norm, self.idf, self.k1, self.dense_tf,
                             self.dense_terms, q_terms, q_off, max_terms, k)

    def scores_tiled(self, q_terms: Tensor, q_off: Tensor, max_terms: int, out: Tensor) -> None:
        """get_scores of a batch written into the tiled matrix ``out[ceil(n_docs / 256), rows >= B, 256]``."""
        ops.bm25_scores_tiled(self.term_off, self.post_doc, self.post_tf, self.norm, self.idf, self.k1,
                              self.dense_tf, self.dense_terms, q_terms, q_off, max_terms, out)

    def scores(self, q_terms: Tensor, q_off: Tensor, max_terms: int) -> Tensor:
        return ops.bm25_scores(self.term_off, self.post_doc, self.post_tf, self.norm, self.idf, self.k1,
                               self.dense_tf, self.dense_terms, q_terms, q_off, max_terms)


def _segment(doc_off: Tensor, doc_tok: Tensor, vocab: int):
    """(term, local doc, tf) triples sorted by (term, doc) + per-term posting counts for one block of documents."""
    n = doc_off.shape[0] - 1
    lens = doc_off[1:] - doc_off[:-1]
    owner = torch.repeat_interleave(torch.arange(n, device=doc_tok.device, dtype=torch.int64), lens)
    key = doc_tok.to(torch.int64) * n + owner
    key = torch.sort(key).values
    pair, tf = torch.unique_consecutive(key, return_counts=True)
    term = pair // n
    doc = (pair - term * n).to(torch.int32)
    counts = torch.bincount(term, minlength=vocab)
    return term, doc, tf.clamp_(max=65535).to(torch.int16), counts, lens.to(torch.int32)


def build_shard(doc_off: Tensor, doc_tok: Tensor, vocab: int, id_base: int = 0, k1: float = 1.5, b: float = 0.75,
                epsilon: float = 0.25) -> SparseShard:
    """CSR from a doc-major token list (doc_off int64 [N+1], doc_tok int32 [total]) on doc_tok's device."""
    if doc_tok.numel() and (int(doc_tok.min()) < 0 or int(doc_tok.max()) >= vocab):
        raise ValueError("token id outside [0, vocab)")
    term, doc, tf, counts, lens = _segment(doc_off.to(doc_tok.device), doc_tok, vocab)
    term_off = torch.zeros(vocab + 1, dtype=torch.int64, device=doc_tok.device)
    torch.cumsum(counts, 0, out=term_off[1:])
    return SparseShard(term_off, doc.contiguous(), tf.contiguous(), lens.contiguous(), counts.to(torch.int32),
                       n_docs=int(lens.shape[0]), vocab=vocab, id_base=id_base, k1=k1, b=b, epsilon=epsilon)


def build_shard_blocked(block_iter, n_docs: int, vocab: int, device, id_base: int = 0, k1: float = 1.5,
                        b: float = 0.75, epsilon: float = 0.25) -> SparseShard:
    """Same result as ``build_shard`` for corpora too large to sort at once.

    ``block_iter`` yields (doc_off, doc_tok) for consecutive blocks of documents.  Every block
    is segmented on its own; blocks are then scattered into the term-major layout in document
    order, which keeps post_doc ascending inside each term.
    """
    blocks: List[tuple] = []
    total_counts = torch.zeros(vocab, dtype=torch.int64, device=device)
    first = 0
    for doc_off, doc_tok in block_iter:
        term, doc, tf, counts, lens = _segment(doc_off.to(device), doc_tok.to(device), vocab)
        blocks.append((term.to(torch.int32), doc, tf, counts, lens, first))
        total_counts += counts
        first += int(lens.shape[0])
    if first != n_docs:
        raise ValueError(f"blocks hold {first} documents, expected {n_docs}")
    term_off = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
    torch.cumsum(total_counts, 0, out=term_off[1:])
    nnz = int(term_off[-1])
    post_doc = torch.empty(nnz, dtype=torch.int32, device=device)
    post_tf = torch.empty(nnz, dtype=torch.int16, device=device)
    doc_len = torch.empty(n_docs, dtype=torch.int32, device=device)
    placed = torch.zeros(vocab, dtype=torch.int64, device=device)
    while blocks:
        term, doc, tf, counts, lens, base = blocks.pop(0)
        t64 = term.to(torch.int64)
        start_in_block = torch.cumsum(counts, 0) - counts
        dst = term_off[:-1][t64] + placed[t64] + (torch.arange(t64.shape[0], device=device) - start_in_block[t64])
        post_doc[dst] = doc + base
        post_tf[dst] = tf
        doc_len[base:base + lens.shape[0]] = lens
        placed += counts
        del term, doc, tf, counts, lens, t64, dst
    return SparseShard(term_off, post_doc, post_tf, doc_len, total_counts.to(torch.int32), n_docs=n_docs, vocab=vocab,
                       id_base=id_base, k1=k1, b=b, epsilon=epsilon)


def grow_vocab(shard: SparseShard, vocab: int) -> None:
    """New terms have no postings in an existing segment: extend its directory in O(V)."""
    extra = vocab - shard.vocab
    if extra <= 0:
        return
    dev = shard.term_off.device
    shard.term_off = torch.cat([shard.term_off, shard.term_off[-1:].expand(extra)]).contiguous()
    shard.df = torch.cat([shard.df, torch.zeros(extra, dtype=shard.df.dtype, device=dev)])
    shard.vocab = vocab


def merge_adjacent(a: SparseShard, b: SparseShard) -> SparseShard:
    """Concatenate two segments of consecutive rows (a first) into one term-major CSR."""
    assert a.vocab == b.vocab and b.id_base == a.id_base + a.n_docs
    dev = a.post_doc.device
    ca, cb = a.term_off[1:] - a.term_off[:-1], b.term_off[1:] - b.term_off[:-1]
    term_off = torch.zeros(a.vocab + 1, dtype=torch.int64, device=dev)
    torch.cumsum(ca + cb, 0, out=term_off[1:])
    post_doc = torch.empty(a.nnz + b.nnz, dtype=torch.int32, device=dev)
    post_tf = torch.empty(a.nnz + b.nnz, dtype=torch.int16, device=dev)
    terms = torch.arange(a.vocab, device=dev)
    ta, tb = torch.repeat_interleave(terms, ca), torch.repeat_interleave(terms, cb)
    dst_a = term_off[:-1][ta] + (torch.arange(a.nnz, device=dev) - a.term_off[:-1][ta])
    dst_b = term_off[:-1][tb] + ca[tb] + (torch.arange(b.nnz, device=dev) - b.term_off[:-1][tb])
    post_doc[dst_a], post_tf[dst_a] = a.post_doc, a.post_tf
    post_doc[dst_b], post_tf[dst_b] = b.post_doc + a.n_docs, b.post_tf
    return SparseShard(term_off, post_doc, post_tf, torch.cat([a.doc_len, b.doc_len]), a.df + b.df,
                       n_docs=a.n_docs + b.n_docs, vocab=a.vocab, id_base=a.id_base, k1=a.k1, b=a.b, epsilon=a.epsilon,
                       use_dense_table=a.use_dense_table)


class SegmentedIndex:
    """Append-only BM25 index: every ``append`` builds a CSR segment for the NEW documents only.

    "Next" row N1 of SURVEY.md section 8f.  The reference rebuilds ``BM25Okapi`` over the whole
    corpus on every ``add_documents`` (rag_uq/streaming_index.py:140-142) because every idf and
    the average length change; here the postings of old documents never move: an append costs
    O(new tokens) for the segment plus O(V + N) to refresh idf / norm on all segments (and the
    dense tf rows only when the set of frequent terms changes).  Segments are row shards on one
    GPU: each is scored with the global statistics and the per-segment lists are merged, exactly
    like the multi-GPU path.  More than ``max_segments`` segments trigger a merge of the two
    smallest neighbours.
    """

    def __init__(self, k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25, max_segments: int = 8):
        self.k1, self.b, self.epsilon, self.max_segments = k1, b, epsilon, max_segments
        self.segments: List[SparseShard] = []
        self.vocab = 0
        self.n_docs = 0
        self.total_len = 0
        self.use_dense_table = True

    def append(self, doc_off: Tensor, doc_tok: Tensor, vocab: int) -> None:
        seg = build_shard(doc_off, doc_tok, vocab, id_base=self.n_docs, k1=self.k1, b=self.b, epsilon=self.epsilon)
        seg.use_dense_table = self.use_dense_table
        seg.bake_impacts = False
        self.vocab = max(self.vocab, vocab)
        self.segments.append(seg)
        self.n_docs += seg.n_docs
        self.total_len += int(seg.doc_len.sum())
        while len(self.segments) > self.max_segments:
            sizes = [self.segments[i].n_docs + self.segments[i + 1].n_docs for i in range(len(self.segments) - 1)]
            i = sizes.index(min(sizes))
            for s in (self.segments[i], self.segments[i + 1]):
                grow_vocab(s, self.vocab)
            self.segments[i:i + 2] = [merge_adjacent(self.segments[i], self.segments[i + 1])]
            self.segments[i].bake_impacts = False
        self.refresh()

    def refresh(self) -> None:
        for s in self.segments:
            grow_vocab(s, self.vocab)
        df = self.segments[0].df.clone()
        for s in self.segments[1:]:
            df += s.df
        for s in self.segments:
            s.finalize(df, self.n_docs, self.total_len)

    @property
    def post_doc(self) -> Tensor:   # device probe used by callers
        return self.segments[0].post_doc

    @property
    def idf(self) -> Tensor:
        return self.segments[0].idf

    def score_topk(self, q_terms: Tensor, q_off: Tensor, max_terms: int, k: int, seed: Optional[Tensor] = None):
        parts = [s.score_topk(q_terms, q_off, max_terms, k, seed) for s in self.segments]
        if len(parts) == 1:
            return parts[0]
        return ops.topk_merge(torch.stack([p[0] for p in parts], 1), torch.stack([p[1] for p in parts], 1), k)

    def scores(self, q_terms: Tensor, q_off: Tensor, max_terms: int) -> Tensor:
        return torch.cat([s.scores(q_terms, q_off, max_terms) for s in self.segments], dim=1)
