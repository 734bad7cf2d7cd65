"""Generate golden vectors from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports ``rag_uq.router`` from /root/reference (read-only, untouched) and stores
inputs + the reference's own outputs for the router path in
``tests/golden/router_golden.npz``.  /root/reference does not exist on the GPU
box, so tests only ever read the .npz.

BM25 known answers (``bm25_known_answers.json``) are NOT produced by reference
code - rank_bm25 is an un-vendored, un-installable dependency - they are the
hand-derived values listed in SURVEY.md section 8(c4), re-derived here from the
published Okapi formula in plain Python floats (no oracle import) so the oracle
is checked against an independent computation.
"""
import json
import math
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference")
from rag_uq.router import RetrievalRouter, RouterConfig  # noqa: E402


def router_cases():
    out = {}
    for tag, hidden in (("h64", 64), ("h32", 32)):
        torch.manual_seed(7)
        router = RetrievalRouter(RouterConfig(hidden_dim=hidden))
        router.eval()
        for key, val in router.state_dict().items():
            out[f"{tag}/state/{key}"] = val.detach().numpy().copy()

        g = torch.Generator().manual_seed(11)
        bm25 = torch.rand(4, 20, generator=g) * 10.0          # run_router_training.py:251 U(0,10)
        dense = torch.rand(4, 20, generator=g)                # :252 U(0,1)
        bm25[0, 3] = 0.0                                      # a pool miss (streaming_index.py:498)
        dense[1, 5] = 0.0
        out[f"{tag}/bm25"] = bm25.numpy().copy()
        out[f"{tag}/dense"] = dense.numpy().copy()

        with torch.no_grad():
            # (1) state right after load_state_dict: stats_initialized False -> batch-wise norm
            out[f"{tag}/gate_batchstat"] = router(bm25, dense).numpy().copy()
            vals, idx = router.hybrid_rerank(bm25, dense, top_k=10)
            out[f"{tag}/rerank_batchstat_vals"] = vals.numpy().copy()
            out[f"{tag}/rerank_batchstat_idx"] = idx.numpy().copy()
            # (2) running statistics armed
            router.bm25_mean.fill_(4.7)
            router.bm25_std.fill_(2.9)
            router.dense_mean.fill_(0.48)
            router.dense_std.fill_(0.31)
            router.stats_initialized = True
            out[f"{tag}/running_stats"] = np.array([4.7, 2.9, 0.48, 0.31], dtype=np.float32)
            out[f"{tag}/gate_running"] = router(bm25, dense).numpy().copy()
            vals, idx = router.hybrid_rerank(bm25, dense, top_k=10)
            out[f"{tag}/rerank_running_vals"] = vals.numpy().copy()
            out[f"{tag}/rerank_running_idx"] = idx.numpy().copy()
            big_b = torch.rand(2, 300, generator=g) * 10.0
            big_d = torch.rand(2, 300, generator=g)
            out[f"{tag}/big_bm25"] = big_b.numpy().copy()
            out[f"{tag}/big_dense"] = big_d.numpy().copy()
            vals, idx = router.hybrid_rerank(big_b, big_d, top_k=500)   # k > P clamps (router.py:202)
            out[f"{tag}/big_rerank_vals"] = vals.numpy().copy()
            out[f"{tag}/big_rerank_idx"] = idx.numpy().copy()

            # (3) MC-Dropout: Dropout active, update_stats=False (no EMA side effect, router.py:114)
            router.train()
            T = 6
            masks, gates = [], []
            for t in range(T):
                torch.manual_seed(1000 + t)
                masks.append(torch.empty(bm25.numel(), hidden).bernoulli_(0.9).numpy().astype(np.uint8))
                torch.manual_seed(1000 + t)
                gates.append(router(bm25, dense, update_stats=False).numpy().copy())
            router.eval()
            out[f"{tag}/mc_masks"] = np.stack(masks)
            out[f"{tag}/mc_gates"] = np.stack(gates)
    return out


def okapi_plain(corpus, query, k1=1.5, b=0.75, eps=0.25):
    """Okapi BM25 with the rank_bm25 conventions, scalar Python only."""
    docs = [t.lower().split() for t in corpus]
    n = len(docs)
    avgdl = sum(len(d) for d in docs) / n
    vocab = []
    for d in docs:
        for w in d:
            if w not in vocab:
                vocab.append(w)
    idf = {}
    for w in vocab:
        nd = sum(1 for d in docs if w in d)
        idf[w] = math.log(n - nd + 0.5) - math.log(nd + 0.5)
    avg_idf = sum(idf.values()) / len(idf)
    idf = {w: (eps * avg_idf if v < 0 else v) for w, v in idf.items()}
    scores = []
    for d in docs:
        s = 0.0
        for q in query.lower().split():
            tf = d.count(q)
            s += idf.get(q, 0.0) * (tf * (k1 + 1) / (tf + k1 * (1 - b + b * len(d) / avgdl)))
        scores.append(s)
    return scores, avgdl, avg_idf, idf


def bm25_cases():
    corpus = [
        "the sky is blue",
        "the sun is bright",
        "the sun in the sky is bright",
        "we can see the shining sun the bright sun",
        "python is a programming language",
        "machine learning uses python",
    ]
    cases = {"corpus": corpus, "k1": 1.5, "b": 0.75, "epsilon": 0.25, "queries": {}}
    for q in ["sun sky", "the sun", "python python language", "zzz", "The SKY", ""]:
        scores, avgdl, avg_idf, idf = okapi_plain(corpus, q)
        cases["queries"][q] = scores
    cases["avgdl"] = avgdl
    cases["average_idf"] = avg_idf
    cases["idf"] = idf
    # SURVEY.md section 8(c4) values, kept verbatim as a cross-check of this script itself
    survey = {
        "avgdl": 5.5, "average_idf": 0.8661886560868407,
        "sun sky": [0.6700158874531926, 0, 0.5235346812893369, 0, 0, 0],
        "the sun": [0.24684132686412555, 0.24684132686412555, 0.2844201557300074, 0.2568214344192789, 0, 0],
        "python python language": [0, 0, 0, 0, 2.5804189055241222, 1.3400317749063853],
    }
    assert abs(avgdl - survey["avgdl"]) < 1e-12 and abs(avg_idf - survey["average_idf"]) < 1e-12
    for q in ("sun sky", "the sun", "python python language"):
        assert np.allclose(cases["queries"][q], survey[q], rtol=1e-12, atol=1e-15), q
    return cases


if __name__ == "__main__":
    np.savez_compressed(HERE / "router_golden.npz", **router_cases())
    with open(HERE / "bm25_known_answers.json", "w") as f:
        json.dump(bm25_cases(), f, indent=1)
    print("wrote", HERE / "router_golden.npz", HERE / "bm25_known_answers.json")
