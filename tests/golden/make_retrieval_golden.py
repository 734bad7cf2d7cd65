"""Golden vectors for the retrieval half of the path, produced by the LIVE reference classes.

    PYTHONHASHSEED=0 PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_retrieval_golden.py

Run in the build container only (/root/reference does not exist on the GPU box; tests read the JSON).

What runs here is the reference's own code, imported from /root/reference and left untouched:

* ``rag_uq.streaming_index.BM25Index`` (:92-225) - ``add_documents`` + ``search`` (tokenise, get_scores,
  ``np.argsort(...)[::-1][:top_k]``, the ``> 0`` filter, row -> doc id).  The one thing it cannot import in
  this image is the third-party ``rank_bm25.BM25Okapi`` (un-vendored, not installable), so the module
  attribute ``BM25Okapi`` is bound to ``oracle.bm25_okapi.OkapiLiteral`` (the restated rank_bm25 0.2.2
  arithmetic, pinned separately by the known-answer vectors).  Everything AROUND that arithmetic is the
  reference's.
* ``rag_uq.streaming_index.HybridRetriever`` (:376-560) - constructed with both indices disabled
  (:405-420), its ``bm25_search`` / ``dense_search`` replaced per case by fixed pools, so
  ``hybrid_search`` (:464-523) and ``get_scores_for_router`` (:525-557) run exactly as written: union,
  0.0 for a missing side, ids without a stored document dropped BEFORE the maxima, ``max(...) or 1``,
  average, stable descending sort, cut, padding with 0.0 / "".

Ties: the reference's order among equal hybrid scores is set-iteration order and among equal BM25 scores
an artefact of introsort; a case whose cut would fall inside a tie group is rejected here (asserted), and
consumers compare tie groups as sets.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import rag_uq.streaming_index as ref  # noqa: E402
from oracle import bm25_okapi  # noqa: E402

ref.BM25Okapi = bm25_okapi.OkapiLiteral   # the un-vendored dependency; see the module docstring


def _no_boundary_tie(scores, cut):
    return cut >= len(scores) or cut == 0 or scores[cut - 1] != scores[cut]


# ------------------------------------------------------------------------------------------ BM25Index.search
WORDS = ("alpha beta gamma delta epsilon zeta eta theta iota kappa lambda mu nu xi omicron pi rho sigma tau upsilon "
         "phi chi psi omega the of and to in is that for it with as was on be at by this had not are but from").split()


def _random_corpus(rng, n_docs, zipf_a):
    p = 1.0 / np.arange(1, len(WORDS) + 1) ** zipf_a
    p /= p.sum()
    docs = []
    for i in range(n_docs):
        n = int(rng.integers(3, 40))
        words = rng.choice(WORDS, size=n, p=p).tolist()
        if i % 7 == 0:
            words = [w.upper() if j % 3 == 0 else w for j, w in enumerate(words)]     # _tokenize lower-cases
        docs.append(("  " if i % 5 == 0 else " ").join(words))                        # and splits on runs of blanks
    return docs


def bm25_index_cases():
    rng = np.random.default_rng(20261018)
    out = []
    corpora = {
        "survey_c4": ["the sky is blue", "the sun is bright", "the sun in the sky is bright",
                      "we can see the shining sun the bright sun", "python is a programming language",
                      "machine learning uses python"],
        "zipf_120": _random_corpus(rng, 120, 1.1),
        "zipf_37_flat": _random_corpus(rng, 37, 0.3),
        "one_doc": ["lonely document with lonely words"],
    }
    queries = {
        "survey_c4": [("sun sky", 10), ("the sun", 2), ("the sun", 4), ("python python language", 10), ("zzz", 10), ("The SKY", 2), ("", 5),
                      ("sun", 10)],
        "zipf_120": [("alpha omega", 10), ("the of and", 50), ("psi", 200), ("tau TAU tau", 7), ("kappa unknownword mu", 200), ("kappa unknownword mu", 1),
                     ("omega", 1), ("chi phi upsilon tau sigma rho pi omicron", 50)],
        "zipf_37_flat": [("beta gamma", 10), ("this had not are but from", 37), ("delta", 100)],
        "one_doc": [("lonely", 5), ("words document", 1), ("absent", 3)],
    }
    for name, texts in corpora.items():
        index = ref.BM25Index()
        docs = [ref.Document(id=f"{name}-{i:03d}", text=t) for i, t in enumerate(texts)]
        assert index.add_documents(docs[: len(docs) // 2 + 1]) == len(docs) // 2 + 1
        assert index.add_documents(docs) == len(docs) - (len(docs) // 2 + 1)      # duplicates skipped (:135), rebuild (:142)
        cases = []
        for query, top_k in queries[name]:
            full = index.bm25.get_scores(index._tokenize(query))
            ranked = sorted(full.tolist(), reverse=True)
            if not _no_boundary_tie(ranked, top_k) and ranked[top_k - 1] > 0:
                raise AssertionError(f"{name!r} / {query!r}: the cut at {top_k} falls inside a tie group; pick another case")
            got = index.search(query, top_k)
            cases.append({"query": query, "top_k": top_k, "result": [[d, s] for d, s in got]})
        out.append({"name": name, "k1": index.k1, "b": index.b, "doc_ids": [d.id for d in docs], "texts": texts,
                    "cases": cases})
    return out


# ------------------------------------------------------------------------------------------ HybridRetriever
def _retriever(doc_ids, bm25_pool, dense_pool):
    r = ref.HybridRetriever()              # HAS_BM25 = HAS_CHROMA = False here: both indices None (:405-420)
    assert r.bm25_index is None and r.dense_index is None
    for d in doc_ids:
        r.documents[d] = ref.Document(id=d, text=f"text of {d}", title=f"title {d}")
    seen = {}

    def bm25_search(query, top_k=20):
        seen["bm25_top_k"] = top_k
        return list(bm25_pool[:top_k])

    def dense_search(query, top_k=20):
        seen["dense_top_k"] = top_k
        return list(dense_pool[:top_k])

    r.bm25_search, r.dense_search = bm25_search, dense_search
    return r, seen


def hybrid_cases():
    rng = np.random.default_rng(424242)
    ids = [f"p{i:04d}" for i in range(400)]

    def pool(n, lo, hi, among=ids, sort=True):
        chosen = rng.choice(among, size=n, replace=False).tolist()
        scores = rng.uniform(lo, hi, size=n)
        scores = np.sort(scores)[::-1] if sort else scores
        return [[c, float(s)] for c, s in zip(chosen, scores)]

    specs = []
    # overlapping pools, the common case (pool 50, top 10)
    shared = rng.choice(ids, size=20, replace=False).tolist()
    b = pool(30, 0.5, 14.0, [i for i in ids if i not in shared]) + [[i, float(s)] for i, s in zip(shared, rng.uniform(0.5, 14, 20))]
    d = pool(30, 0.05, 0.9, [i for i in ids if i not in shared]) + [[i, float(s)] for i, s in zip(shared, rng.uniform(0.05, 0.9, 20))]
    b.sort(key=lambda r: -r[1]); d.sort(key=lambda r: -r[1])
    specs.append(("overlap_pool50_top10", ids, b, d, 10, 50, 20))
    specs.append(("overlap_pool50_top100", ids, b, d, 100, 50, 10))
    # disjoint pools
    specs.append(("disjoint", ids, pool(50, 1.0, 9.0, ids[:200]), pool(50, 0.1, 0.8, ids[200:]), 10, 50, 20))
    # the BM25 side returns nothing (every query word out of vocabulary): max(bm25) = 0 -> "or 1"
    specs.append(("bm25_empty", ids, [], pool(50, 0.1, 0.8), 10, 50, 10))
    # the dense side returns nothing
    specs.append(("dense_empty", ids, pool(12, 1.0, 9.0), [], 10, 50, 20))
    # both empty
    specs.append(("both_empty", ids, [], [], 10, 50, 5))
    # dense scores all zero (the reference's failed-embedding default gives cosine 0): max = 0 -> "or 1"
    specs.append(("dense_all_zero", ids, pool(8, 1.0, 9.0), [[i, 0.0] for i in rng.choice(ids, 6, replace=False).tolist()], 20, 50, 20))
    # negative cosines: mixed sign, and ALL negative (the maximum is negative, the division flips the order)
    specs.append(("dense_mixed_sign", ids, pool(20, 1.0, 9.0), pool(20, -0.4, 0.6), 10, 50, 10))
    specs.append(("dense_all_negative", ids, [], pool(15, -0.9, -0.05), 10, 50, 10))
    specs.append(("dense_all_negative_with_bm25", ids, pool(10, 0.5, 7.0), pool(15, -0.9, -0.05), 12, 50, 20))
    # ids the retriever holds no document for are dropped BEFORE the maxima are taken (:494-496);
    # the ghosts carry the largest scores on both sides so the maxima change
    ghosts_b = [["ghost-b1", 99.0], ["ghost-b2", 50.0]]
    ghosts_d = [["ghost-d1", 0.999], ["ghost-b1", 0.99]]
    specs.append(("ghost_ids", ids, ghosts_b + pool(20, 1.0, 9.0), ghosts_d + pool(20, 0.1, 0.8), 10, 50, 20))
    # fewer hits than requested: padding with 0.0 / ""
    specs.append(("short", ids, pool(3, 1.0, 9.0), pool(2, 0.1, 0.8), 10, 50, 20))
    # pool size smaller than what the retrievers could return (the cut happens inside the stubs, like search(top_k))
    specs.append(("pool_cut_5", ids, pool(30, 1.0, 9.0), pool(30, 0.1, 0.8), 8, 5, 8))
    # random sweep
    for j in range(12):
        nb, nd = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        specs.append((f"random_{j:02d}", ids, pool(nb, 0.2, 20.0), pool(nd, -0.2, 1.0), int(rng.integers(1, 40)),
                      int(rng.integers(1, 70)), int(rng.integers(1, 30))))

    out = []
    for name, doc_ids, bpool, dpool, top_k, pool_size, num_passages in specs:
        r, seen = _retriever(doc_ids, [tuple(x) for x in bpool], [tuple(x) for x in dpool])
        res = r.hybrid_search("ignored", top_k=top_k, retrieval_pool_size=pool_size)
        assert seen == {"bm25_top_k": pool_size, "dense_top_k": pool_size}
        everything = r.hybrid_search("ignored", top_k=10 ** 6, retrieval_pool_size=pool_size)
        hs = [x.hybrid_score for x in everything]
        assert _no_boundary_tie(hs, top_k), f"{name}: cut inside a tie group"
        every50 = r.hybrid_search("ignored", top_k=10 ** 6)
        while not _no_boundary_tie([x.hybrid_score for x in every50], num_passages):
            num_passages += 1                                       # keep the router cut outside a tie group
        sb, sd, sids, stexts = r.get_scores_for_router("ignored", num_passages=num_passages)
        assert seen == {"bm25_top_k": 50, "dense_top_k": 50}         # get_scores_for_router always pools 50 (:537)
        out.append({
            "name": name, "n_documents": len(doc_ids), "bm25_pool": bpool, "dense_pool": dpool,
            "top_k": top_k, "retrieval_pool_size": pool_size,
            "hybrid_search": [[x.doc_id, x.bm25_score, x.dense_score, x.hybrid_score, x.text, x.title] for x in res],
            "num_passages": num_passages,
            "scores_for_router": {"bm25": sb, "dense": sd, "ids": sids, "texts": stexts},
        })
    return {"document_ids": ids, "cases": out}


if __name__ == "__main__":
    payload = {"generator": "tests/golden/make_retrieval_golden.py (live /root/reference classes; BM25Okapi bound to "
                            "oracle.bm25_okapi.OkapiLiteral, bm25_search/dense_search stubbed per case)",
               "bm25_index_search": bm25_index_cases(), "hybrid": hybrid_cases()}
    path = HERE / "retrieval_golden.json"
    path.write_text(json.dumps(payload, indent=1))
    print("wrote", path, path.stat().st_size, "bytes")
