"""String-level drop-in for the reference's retrieval classes (rag_uq/streaming_index.py).

Same class names, constructor arguments, method signatures, return types and soft-failure
behaviour as ``Document`` (:54-77), ``RetrievalResult`` (:80-89), ``BM25Index`` (:92-225),
``DenseIndex`` (:228-373), ``HybridRetriever`` (:376-560) and ``StreamingIndex`` (:563-686);
all scoring happens on the B200 through ``engine.HybridEngine``.  Differences, all additive:

* no rank_bm25 / chromadb needed: the sparse index is a CSR in HBM, the dense index an exact
  bf16 matrix (the reference's HNSW is approximate);
* ``embed_fn`` (texts -> [n, dim] array) replaces the per-text Ollama HTTP call
  (:267-288); without it the class tries ``ollama`` and then the reference's sha256
  pseudo-embedding (:269-273), exactly in that order;
* ``*_batch`` methods take whole query batches;
* the dense side persists under ``persist_directory`` as raw arrays instead of a ChromaDB store
  (``<persist_directory>/<collection_name>/rows.bf16 + meta.jsonl + header.json``, appended on every add and
  reloaded by the constructor, like the reference's PersistentClient at :254-263), so an index resumed from
  the ``StreamingIndex`` checkpoint has its dense rows back; ``persist_directory=None`` keeps it in memory only.
Tie order, which the reference leaves to introsort / set iteration, is fixed: higher score
first, then the document added earlier.
"""
from __future__ import annotations

import hashlib
import itertools
import json
import logging
import os
import pickle
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import HybridEngine
from .sparse import SegmentedIndex

logger = logging.getLogger(__name__)


@dataclass
class Document:
    """A document for indexing (streaming_index.py:54-77)."""
    id: str
    text: str
    title: Optional[str] = None
    metadata: Optional[Dict[str, Any]] = None

    def to_dict(self) -> Dict[str, Any]:
        return {"id": self.id, "text": self.text, "title": self.title or "", "metadata": self.metadata or {}}

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "Document":
        return cls(id=data["id"], text=data["text"], title=data.get("title"), metadata=data.get("metadata"))


@dataclass
class RetrievalResult:
    """Result from hybrid retrieval (streaming_index.py:80-89)."""
    doc_id: str
    text: str
    bm25_score: float
    dense_score: float
    hybrid_score: Optional[float] = None
    title: Optional[str] = None
    metadata: Optional[Dict[str, Any]] = None


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("rag_uq_b200 needs a CUDA device (B200, sm_100); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ==========================================================================================
class BM25Index:
    """BM25 sparse index; scoring follows rank_bm25.BM25Okapi (k1, b, epsilon = 0.25)."""

    def __init__(self, persist_path: Optional[str] = None, k1: float = 1.5, b: float = 0.75, device=None):
        self.persist_path = Path(persist_path) if persist_path else None
        self.k1, self.b = k1, b
        self.device = device
        self.documents: Dict[str, Document] = {}
        self.doc_ids: List[str] = []
        self.tokenized_corpus: List[List[str]] = []
        self.vocab: Dict[str, int] = {}
        self._doc_terms: List[np.ndarray] = []
        self.bm25: Optional[SegmentedIndex] = None   # the reference keeps its BM25Okapi here
        self._indexed = 0                             # documents already in a GPU segment
        self._stale = False
        if self.persist_path and self.persist_path.exists():
            self._load()

    def _tokenize(self, text: str) -> List[str]:
        return text.lower().split()

    def _intern(self, tokens: Sequence[str]) -> np.ndarray:
        vocab = self.vocab
        return np.fromiter((vocab.setdefault(t, len(vocab)) for t in tokens), dtype=np.int32, count=len(tokens))

    def add_documents(self, documents: List[Document]) -> int:
        added = 0
        for doc in documents:
            if doc.id in self.documents:
                continue
            self.documents[doc.id] = doc
            self.doc_ids.append(doc.id)
            tokens = self._tokenize(doc.text)
            self.tokenized_corpus.append(tokens)
            self._doc_terms.append(self._intern(tokens))
            added += 1
        if added and self.tokenized_corpus:
            self._stale = True      # statistics of EVERY posting change; rebuilt lazily on the GPU
            logger.info(f"Added {added} documents to BM25 index. Total: {len(self.doc_ids)}")
        if self.persist_path:
            self._save()
        return added

    def _ensure_built(self) -> None:
        """Index the documents added since the last search as ONE new segment (incremental ingest)."""
        if not self._stale and self.bm25 is not None:
            return
        dev = torch.device(self.device) if self.device is not None else _default_device()
        fresh = self._doc_terms[self._indexed:]
        lens = np.fromiter((len(t) for t in fresh), dtype=np.int64, count=len(fresh))
        doc_off = np.concatenate([[0], np.cumsum(lens)])
        doc_tok = np.concatenate(fresh) if len(fresh) else np.zeros(0, np.int32)
        if self.bm25 is None:
            self.bm25 = SegmentedIndex(k1=self.k1, b=self.b)
        self.bm25.append(torch.from_numpy(doc_off).to(dev), torch.from_numpy(doc_tok).to(dev),
                         vocab=max(len(self.vocab), 1))
        self._indexed = len(self._doc_terms)
        self._stale = False

    def encode_queries(self, queries: Sequence[str]):
        """Batched tokeniser + vocabulary lookup: -> (q_terms int32, q_off int32, max_terms) on the index device;
        unknown words become -1 (they contribute nothing, like ``idf.get(q) or 0`` in rank_bm25).

        Tokenisation is the reference's (``text.lower().split()``, :118-120) applied per query; the id lookup runs as
        ONE C-level pass over all tokens of the batch (``map`` over the interned-vocabulary dict straight into a numpy
        buffer), not a Python loop per token."""
        rows = [q.lower().split() for q in queries]
        lens = np.fromiter(map(len, rows), dtype=np.int64, count=len(rows))
        longest = int(lens.max()) if len(rows) else 0
        if longest > _lib.MAX_QUERY_TERMS:
            raise ValueError(f"a query has {longest} tokens; the BM25 kernel accepts at most {_lib.MAX_QUERY_TERMS}")
        total = int(lens.sum())
        flat = np.fromiter(map(self.vocab.get, itertools.chain.from_iterable(rows), itertools.repeat(-1)), dtype=np.int32,
                           count=total)
        off = np.zeros(len(rows) + 1, dtype=np.int32)
        np.cumsum(lens, out=off[1:])
        dev = self.bm25.post_doc.device if self.bm25 is not None else _default_device()
        if flat.size == 0:
            flat = np.full(1, -1, dtype=np.int32)
        return torch.from_numpy(flat).to(dev), torch.from_numpy(off).to(dev), max(longest, 1)

    def search_batch(self, queries: Sequence[str], top_k: int = 10):
        """-> (scores [B,k] fp32, rows [B,k] int32) on the device; row -1 pads."""
        self._ensure_built()
        q_terms, q_off, max_terms = self.encode_queries(queries)
        return self.bm25.score_topk(q_terms, q_off, max_terms, top_k)

    def search(self, query: str, top_k: int = 10) -> List[Tuple[str, float]]:
        if not self.doc_ids or (self.bm25 is None and not self._stale):
            return []
        score, rows = self.search_batch([query], min(top_k, _lib.MAX_TOPK))
        score, rows = score[0].tolist(), rows[0].tolist()
        return [(self.doc_ids[r], float(s)) for s, r in zip(score, rows) if r >= 0]

    def get_document(self, doc_id: str) -> Optional[Document]:
        return self.documents.get(doc_id)

    def _save(self) -> None:
        """Same pickle schema as the reference (streaming_index.py:192-201), so either side can load it."""
        if self.persist_path is None:
            return
        self.persist_path.parent.mkdir(parents=True, exist_ok=True)
        payload = {"documents": {k: v.to_dict() for k, v in self.documents.items()}, "doc_ids": self.doc_ids,
                   "tokenized_corpus": self.tokenized_corpus, "k1": self.k1, "b": self.b}
        with open(self.persist_path, "wb") as fh:
            pickle.dump(payload, fh)

    def _load(self) -> None:
        if self.persist_path is None or not self.persist_path.exists():
            return
        with open(self.persist_path, "rb") as fh:
            payload = pickle.load(fh)
        self.documents = {k: Document.from_dict(v) for k, v in payload["documents"].items()}
        self.doc_ids = payload["doc_ids"]
        self.tokenized_corpus = payload["tokenized_corpus"]
        self.k1, self.b = payload["k1"], payload["b"]
        self.vocab, self._doc_terms, self._indexed, self.bm25 = {}, [], 0, None
        for tokens in self.tokenized_corpus:
            self._doc_terms.append(self._intern(tokens))
        self._stale = bool(self.tokenized_corpus)
        logger.info(f"Loaded BM25 index with {len(self.doc_ids)} documents")

    def __len__(self) -> int:
        return len(self.doc_ids)


# ==========================================================================================
class DenseIndex:
    """Exact cosine index over bf16 unit rows in HBM (the reference asks ChromaDB's HNSW)."""

    def __init__(self, collection_name: str = "rag_documents", persist_directory: Optional[str] = "./data/chroma_db",
                 embedding_model: str = "nomic-embed-text", chroma_host: Optional[str] = None, chroma_port: int = 8000,
                 embed_fn: Optional[Callable[[List[str]], Any]] = None, device=None, mma_variant: int = 3):
        self.collection_name = collection_name
        self.persist_directory = persist_directory
        self.embedding_model = embedding_model
        self.embed_fn = embed_fn
        self.device = device
        self.mma_variant = mma_variant
        self.ids: List[str] = []
        self.texts: List[str] = []
        self.metadatas: List[Dict[str, Any]] = []
        self._id_set: set = set()
        self._rows: Optional[torch.Tensor] = None   # bf16 [capacity, dim_padded]
        self._count = 0
        self.dim: Optional[int] = None
        if chroma_host:
            logger.warning("chroma_host=%s ignored: rag_uq_b200 keeps the dense rows in HBM and persists them under "
                           "persist_directory, it does not talk to a ChromaDB server", chroma_host)
        self._store = Path(persist_directory) / collection_name if persist_directory else None
        if self._store is not None and (self._store / "header.json").exists():
            self._load()
        logger.info(f"Initialized DenseIndex with collection '{collection_name}'")

    # -- persistence (stands in for chromadb.PersistentClient, streaming_index.py:254-263) --------------
    def _persist(self, rows: torch.Tensor, ids: Sequence[str], texts: Sequence[str], metadatas: Sequence[Dict[str, Any]]) -> None:
        """Append the new rows (raw little-endian bf16, padded width) and one JSON line per document."""
        if self._store is None:
            return
        self._store.mkdir(parents=True, exist_ok=True)
        header = self._store / "header.json"
        if not header.exists():
            header.write_text(json.dumps({"format": "rag_uq_b200 dense rows v1", "dim": self.dim, "padded_dim": int(rows.shape[1]),
                                          "dtype": "bfloat16", "embedding_model": self.embedding_model}))
        with open(self._store / "rows.bf16", "ab") as fh:
            fh.write(rows.cpu().view(torch.int16).numpy().tobytes())
        with open(self._store / "meta.jsonl", "a") as fh:
            for i, t, m in zip(ids, texts, metadatas):
                fh.write(json.dumps({"id": i, "text": t, "metadata": m}) + "\n")

    def _load(self) -> None:
        head = json.loads((self._store / "header.json").read_text())
        if head.get("format") != "rag_uq_b200 dense rows v1":
            raise ValueError(f"{self._store}: not a rag_uq_b200 dense store")
        meta_path, rows_path = self._store / "meta.jsonl", self._store / "rows.bf16"
        metas = [json.loads(line) for line in meta_path.read_text().splitlines() if line.strip()] if meta_path.exists() else []
        width = int(head["padded_dim"])
        raw = np.fromfile(rows_path, dtype=np.int16) if rows_path.exists() else np.zeros(0, np.int16)
        n = min(len(metas), raw.size // width)          # a crash between the two appends leaves a ragged tail: drop it
        if n < len(metas) or n * width < raw.size:
            logger.warning(f"{self._store}: dropping an incomplete tail ({len(metas)} metadata lines, {raw.size // width} rows)")
        self.dim = int(head["dim"])
        if n:
            dev = torch.device(self.device) if self.device is not None else _default_device()
            rows = torch.from_numpy(raw[:n * width].reshape(n, width).copy()).view(torch.bfloat16).to(dev)
            self._rows, self._count = rows, n
            self.ids = [m["id"] for m in metas[:n]]
            self.texts = [m["text"] for m in metas[:n]]
            self.metadatas = [m.get("metadata") or {} for m in metas[:n]]
            self._id_set = set(self.ids)
        logger.info(f"Loaded dense index with {n} rows from {self._store}")

    # -- embeddings --------------------------------------------------------------------------
    def _get_embedding(self, text: str) -> List[float]:
        if self.embed_fn is not None:
            return list(np.asarray(self.embed_fn([text]), dtype=np.float32)[0])
        try:
            import ollama  # noqa: WPS433 - optional, exactly like the reference
        except ImportError:
            digest = hashlib.sha256(text.encode()).digest()       # reference test fallback (:269-273)
            return [float(b) / 255.0 for b in digest][:384]
        try:
            return ollama.embeddings(model=self.embedding_model, prompt=text)["embedding"]
        except Exception as exc:
            logger.error(f"Embedding failed: {exc}")
            return [0.0] * 768                                    # reference failure default (:281-284)

    def _get_embeddings_batch(self, texts: List[str]) -> np.ndarray:
        if self.embed_fn is not None:
            return np.asarray(self.embed_fn(list(texts)), dtype=np.float32)
        return np.asarray([self._get_embedding(t) for t in texts], dtype=np.float32)

    def _to_rows(self, emb) -> torch.Tensor:
        """fp32 [n, d] -> unit-normalised, zero-padded to a multiple of 64 columns, bf16, on device."""
        dev = torch.device(self.device) if self.device is not None else _default_device()
        x = torch.as_tensor(np.asarray(emb, dtype=np.float32) if not torch.is_tensor(emb) else emb)
        x = x.to(dev, torch.float32)
        if self.dim is None:
            self.dim = int(x.shape[1])
        if x.shape[1] != self.dim:
            raise ValueError(f"embedding width {x.shape[1]} differs from the index width {self.dim}")
        x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-30)     # cosine space (:262)
        pad = (-self.dim) % 64
        if pad:
            x = torch.nn.functional.pad(x, (0, pad))
        return x.to(torch.bfloat16).contiguous()

    # -- build -------------------------------------------------------------------------------
    def add_embeddings(self, ids: Sequence[str], embeddings, texts: Optional[Sequence[str]] = None,
                       metadatas: Optional[Sequence[Dict[str, Any]]] = None) -> int:
        rows = self._to_rows(embeddings)
        n = rows.shape[0]
        if self._rows is None or self._count + n > self._rows.shape[0]:
            cap = max(2 * (self._count + n), 1024)
            grown = torch.empty((cap, rows.shape[1]), dtype=torch.bfloat16, device=rows.device)
            if self._rows is not None:
                grown[:self._count] = self._rows[:self._count]
            self._rows = grown
        self._rows[self._count:self._count + n] = rows
        self._count += n
        texts = list(texts) if texts is not None else [""] * n
        metadatas = list(metadatas) if metadatas is not None else [{} for _ in range(n)]
        self.ids.extend(ids)
        self._id_set.update(ids)
        self.texts.extend(texts)
        self.metadatas.extend(metadatas)
        self._persist(rows, list(ids), texts, metadatas)
        return n

    def add_documents(self, documents: List[Document], batch_size: int = 100) -> int:
        fresh = [d for d in documents if d.id not in self._id_set]
        if not fresh:
            logger.info("No new documents to add")
            return 0
        total = 0
        for start in range(0, len(fresh), batch_size):
            batch = fresh[start:start + batch_size]
            emb = self._get_embeddings_batch([d.text for d in batch])
            total += self.add_embeddings([d.id for d in batch], emb, [d.text for d in batch],
                                         [{"title": d.title or "", **(d.metadata or {})} for d in batch])
            logger.info(f"Indexed batch {start // batch_size + 1}, total: {total}/{len(fresh)}")
        return total

    @property
    def matrix(self) -> Optional[torch.Tensor]:
        return None if self._rows is None else self._rows[:self._count]

    # -- search ------------------------------------------------------------------------------
    def search_batch(self, query_embeddings, top_k: int = 10):
        """-> (cosine [B,k] fp32, rows [B,k] int32)."""
        if self._count == 0:
            raise ValueError("dense index is empty")
        q = self._to_rows(query_embeddings)
        eng = HybridEngine(None, self.matrix, 0, mma_variant=self.mma_variant)
        return eng.dense_local_topk(q, top_k)

    def search(self, query: str, top_k: int = 10) -> List[Tuple[str, float, str]]:
        if self._count == 0:
            return []
        k = min(top_k, self._count, _lib.MMA_MAX_TOPK)
        score, rows = self.search_batch(np.asarray([self._get_embedding(query)], dtype=np.float32), k)
        return [(self.ids[r], float(s), self.texts[r]) for s, r in zip(score[0].tolist(), rows[0].tolist()) if r >= 0]

    def __len__(self) -> int:
        return self._count


# ==========================================================================================
class HybridRetriever:
    """BM25 + dense retrieval with both scores per passage (streaming_index.py:376-560)."""

    def __init__(self, bm25_persist_path: Optional[str] = "./data/bm25_index.pkl",
                 chroma_persist_path: Optional[str] = "./data/chroma_db", chroma_host: Optional[str] = None,
                 embedding_model: str = "nomic-embed-text", embed_fn: Optional[Callable] = None, device=None,
                 mma_variant: int = 3):
        self.bm25_index = BM25Index(persist_path=bm25_persist_path, device=device)
        self.dense_index = DenseIndex(persist_directory=chroma_persist_path,
                                      chroma_host=chroma_host or os.environ.get("CHROMA_HOST"),
                                      embedding_model=embedding_model, embed_fn=embed_fn, device=device,
                                      mma_variant=mma_variant)
        self.documents: Dict[str, Document] = {}
        self._order: Dict[str, int] = {}
        # The reference keeps ``self.documents`` in memory only (:422-423): after a restart its persisted indices
        # still answer bm25_search / dense_search but hybrid_search drops every hit (:494-496) and returns [].
        # Restoring the document store from the reloaded BM25 pickle is additive.
        for doc_id in self.bm25_index.doc_ids:
            self._order[doc_id] = len(self._order)
            self.documents[doc_id] = self.bm25_index.documents[doc_id]

    def add_documents(self, documents: List[Document], batch_size: int = 100) -> Dict[str, int]:
        for doc in documents:
            if doc.id not in self._order:
                self._order[doc.id] = len(self._order)
            self.documents[doc.id] = doc
        stats = {"bm25_added": self.bm25_index.add_documents(documents),
                 "dense_added": self.dense_index.add_documents(documents, batch_size)}
        stats["total_documents"] = len(self.documents)
        return stats

    def bm25_search(self, query: str, top_k: int = 20) -> List[Tuple[str, float]]:
        return self.bm25_index.search(query, top_k)

    def dense_search(self, query: str, top_k: int = 20) -> List[Tuple[str, float]]:
        return [(doc_id, score) for doc_id, score, _ in self.dense_index.search(query, top_k)]

    def hybrid_search_batch(self, queries: Sequence[str], query_embeddings=None, top_k: int = 10,
                            retrieval_pool_size: int = 50):
        """Device-level hybrid search: -> (bm25 rows, dense rows are unified by DOCUMENT ID on the host)."""
        n = len(queries)
        pool = min(retrieval_pool_size, _lib.MMA_MAX_TOPK)
        bm = self.bm25_index.search_batch(queries, pool) if len(self.bm25_index) else None
        if query_embeddings is None:
            query_embeddings = self.dense_index._get_embeddings_batch(list(queries))
        de = self.dense_index.search_batch(query_embeddings, min(pool, len(self.dense_index))) \
            if len(self.dense_index) else None
        return n, bm, de

    def hybrid_search(self, query: str, top_k: int = 10, retrieval_pool_size: int = 50) -> List[RetrievalResult]:
        return self.hybrid_search_many([query], None, top_k, retrieval_pool_size)[0]

    def _row_numbers(self, side: str, ids_of_row: List[str], dev) -> torch.Tensor:
        """int32 [rows of one index] on the device: the retriever's number of the document in each index row, -1
        when the retriever holds no document for it.  Rebuilt only when the index or the document store grew."""
        cache = self.__dict__.setdefault("_row_number_cache", {})
        key = (len(ids_of_row), len(self._order), len(self.documents), str(dev))
        hit = cache.get(side)
        if hit is None or hit[0] != key or hit[2] is not ids_of_row:
            order, docs = self._order, self.documents
            table = np.fromiter((order[d] if d in docs else -1 for d in ids_of_row), dtype=np.int32, count=len(ids_of_row))
            hit = (key, torch.from_numpy(table).to(dev) if len(ids_of_row) else torch.zeros(1, dtype=torch.int32, device=dev),
                   ids_of_row)
            cache[side] = hit
        return hit[1]

    def hybrid_search_many(self, queries: Sequence[str], query_embeddings=None, top_k: int = 10,
                           retrieval_pool_size: int = 50) -> List[List[RetrievalResult]]:
        """``hybrid_search`` for several queries with one pair of kernel launches.

        The two indices number their rows independently (a document can be missing from one of
        them), so the pools are joined on document id here and fused by the library's fusion
        kernel on a common numbering.
        """
        from . import ops
        n, bm, de = self.hybrid_search_batch(queries, query_embeddings, top_k, retrieval_pool_size)
        if bm is None and de is None:
            return [[] for _ in range(n)]
        pool = (bm[0] if bm is not None else de[0]).shape[1]
        dev = (bm[0] if bm is not None else de[0]).device
        # common numbering = insertion order into this retriever; documents the retriever holds no
        # text for are dropped BEFORE fusion, as the reference does (:494-496)
        names = self.__dict__.get("_names")
        if names is None or len(names) != len(self._order):
            names = self.__dict__["_names"] = sorted(self._order, key=self._order.get)     # number -> document id

        def pools(pair, ids_of_row, side, width):
            """Index rows -> retriever numbering on the device: one gather through a cached row -> number table
            (-1 = the retriever holds no document for that row), no Python loop over B x pool."""
            score = torch.zeros((n, width), dtype=torch.float32, device=dev)
            ident = torch.full((n, width), -1, dtype=torch.int32, device=dev)
            if pair is not None:
                table = self._row_numbers(side, ids_of_row, dev)
                rows = pair[1].to(torch.int64)
                mapped = torch.where(rows >= 0, table[rows.clamp(min=0)], torch.full_like(rows, -1, dtype=torch.int32))
                w = pair[0].shape[1]
                ident[:, :w] = mapped
                score[:, :w] = torch.where(mapped >= 0, pair[0], torch.zeros_like(pair[0]))
            return score, ident

        bs, bi = pools(bm, self.bm25_index.doc_ids, "bm25", pool)
        ds, di = pools(de, self.dense_index.ids, "dense", pool)
        k = min(top_k, 2 * pool)
        ids, sb, sd, sh = (t.cpu().tolist() for t in ops.hybrid_fuse_topk(bs, bi, ds, di, k))
        out: List[List[RetrievalResult]] = []
        for qi in range(n):
            rows = []
            for j in range(k):
                if ids[qi][j] < 0:
                    continue
                doc = self.documents[names[ids[qi][j]]]
                rows.append(RetrievalResult(doc_id=doc.id, text=doc.text, bm25_score=sb[qi][j], dense_score=sd[qi][j],
                                            hybrid_score=sh[qi][j], title=doc.title, metadata=doc.metadata))
            out.append(rows)
        return out

    def get_scores_for_router(self, query: str, num_passages: int = 20):
        results = self.hybrid_search(query, top_k=num_passages)
        bm25 = [r.bm25_score for r in results]
        dense = [r.dense_score for r in results]
        ids = [r.doc_id for r in results]
        texts = [r.text for r in results]
        missing = num_passages - len(results)
        if missing > 0:
            bm25 += [0.0] * missing
            dense += [0.0] * missing
            ids += [""] * missing
            texts += [""] * missing
        return bm25, dense, ids, texts

    def __len__(self) -> int:
        return len(self.documents)


# ==========================================================================================
class StreamingIndex:
    """Resumable JSONL -> index streaming (streaming_index.py:563-686), same checkpoint file."""

    def __init__(self, retriever: HybridRetriever, checkpoint_path: str = "./data/index_checkpoint.json",
                 batch_size: int = 100):
        self.retriever = retriever
        self.checkpoint_path = Path(checkpoint_path)
        self.batch_size = batch_size
        self.progress = self._load_checkpoint()
        sides = [getattr(retriever, name, None) for name in ("bm25_index", "dense_index")]
        have = min(len(side) for side in sides) if all(side is not None for side in sides) else None
        if have is not None and self.progress["total_indexed"] > have:
            # e.g. the dense store was kept in memory only (persist_directory=None) or deleted: skipping the
            # checkpointed lines would leave those documents without postings / dense rows for ever (re-adding
            # is idempotent: documents an index already holds are skipped)
            logger.error(f"checkpoint {self.checkpoint_path} says {self.progress['total_indexed']} documents are indexed "
                         f"but the retriever holds {have}: ignoring the checkpoint and re-indexing from the first line")
            self.progress = {"last_offset": 0, "total_indexed": 0, "files_completed": []}

    def _load_checkpoint(self) -> Dict[str, Any]:
        if self.checkpoint_path.exists():
            text = self.checkpoint_path.read_text().strip()
            if text:
                return json.loads(text)
        return {"last_offset": 0, "total_indexed": 0, "files_completed": []}

    def _save_checkpoint(self) -> None:
        self.checkpoint_path.parent.mkdir(parents=True, exist_ok=True)
        self.checkpoint_path.write_text(json.dumps(self.progress))

    def _commit(self, batch: List[Document], offset: int) -> int:
        self.retriever.add_documents(batch)
        self.progress["last_offset"] = offset
        self.progress["total_indexed"] += len(batch)
        self._save_checkpoint()
        logger.info(f"Indexed batch: {len(batch)} docs, total: {self.progress['total_indexed']}")
        return len(batch)

    def stream_from_jsonl(self, jsonl_path: str, resume: bool = True) -> Iterator[int]:
        path = Path(jsonl_path)
        if not path.exists():
            raise FileNotFoundError(f"Corpus file not found: {jsonl_path}")
        skip = self.progress["last_offset"] if resume else 0
        pending: List[Document] = []
        with open(path) as fh:
            for offset, line in enumerate(fh, start=1):
                if offset <= skip:
                    continue
                try:
                    rec = json.loads(line.strip())
                    pending.append(Document(id=rec["id"], text=rec["text"], title=rec.get("title"),
                                            metadata=rec.get("metadata")))
                except (json.JSONDecodeError, KeyError) as exc:
                    logger.warning(f"Skipping invalid line at offset {offset - 1}: {exc}")
                if len(pending) >= self.batch_size:
                    yield self._commit(pending, offset)
                    pending = []
            if pending:
                yield self._commit(pending, offset)
        if jsonl_path not in self.progress["files_completed"]:
            self.progress["files_completed"].append(jsonl_path)
            self._save_checkpoint()
        logger.info(f"Completed indexing {jsonl_path}")

    def get_progress(self) -> Dict[str, Any]:
        return {**self.progress, "retriever_size": len(self.retriever)}
