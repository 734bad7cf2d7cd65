"""CPU oracle for the rag_uq retrieval-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline - never as the thing shipped.

What is restated here and how it is pinned
------------------------------------------
* ``bm25_okapi``  - rank_bm25 0.2.2 ``BM25Okapi`` (PyPI dependency of the
  reference, ``requirements.txt:8``; NOT vendored under /root/reference and not
  installable in this image) + ``BM25Index.search``
  (``rag_uq/streaming_index.py:150-179``).  ``BM25Index.search`` is **pinned**: the
  live class runs in ``tests/golden/make_retrieval_golden.py`` (its missing third-party
  ``BM25Okapi`` bound to ``OkapiLiteral``) -> ``tests/golden/retrieval_golden.json``.
  The rank_bm25 ARITHMETIC stays **parity unpinned** (package not installable, no
  reference fixture): anchored on the published algorithm and the hand-derived
  known-answer vectors in ``tests/golden/bm25_known_answers.json``.
* ``dense``       - exact cosine scoring that stands in for ChromaDB's
  approximate HNSW (``rag_uq/streaming_index.py:338-370``).  **Parity
  unpinned** (chromadb not installable; no reference test).
* ``fusion``      - ``HybridRetriever.hybrid_search`` /
  ``get_scores_for_router`` (``rag_uq/streaming_index.py:464-557``).  **Pinned**:
  the live ``HybridRetriever`` runs in ``make_retrieval_golden.py`` with stubbed
  pools (disjoint pools, an empty / all-zero side, negative and all-negative
  cosines, ids without a stored document, short pools, pool cuts) and the oracle
  reproduces its output bit for bit.
* ``router``      - ``RetrievalRouter`` forward / hybrid_rerank / MC-Dropout
  (``rag_uq/router.py:100-202``).  **Pinned**: checked bit-for-bit against the
  live reference module imported from /root/reference; golden vectors are
  committed under ``tests/golden/router_golden.npz`` by
  ``tests/golden/make_golden.py``.
* ``hnsw``        - an HNSW index (Malkov & Yashunin; hnswlib conventions, ChromaDB's
  default parameters) restated to REPORT the recall@k of the reference's approximate
  dense path against the exact search built here.  **Parity unpinned** (no chromadb /
  hnswlib in the image, no fixture of their output).
* ``philox``      - Philox4x32-10 with the curand counter layout, used to check
  the in-kernel dropout masks bit-exactly.
"""
