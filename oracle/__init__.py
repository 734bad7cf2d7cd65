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
  (``rag_uq/streaming_index.py:150-179``).  **Parity unpinned**: the reference
  has no test, fixture or golden vector touching ``streaming_index.py``; the
  restatement is anchored on the published rank_bm25 algorithm and on the
  hand-derived known-answer vectors in ``tests/golden/bm25_known_answers.json``.
* ``dense``       - exact cosine scoring that stands in for ChromaDB's
  approximate HNSW (``rag_uq/streaming_index.py:338-370``).  **Parity
  unpinned** (chromadb not installable; no reference test).
* ``fusion``      - ``HybridRetriever.hybrid_search`` /
  ``get_scores_for_router`` (``rag_uq/streaming_index.py:464-557``).  **Parity
  unpinned** (no reference test).
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
