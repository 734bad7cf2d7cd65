#!/usr/bin/env python
"""Small shapes of the new code paths (BM25 window mode + fp16 bound pass, fused full-fusion epilogue, bit-mask dense
epilogue) for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`.
On the round-1 GPU pool compute-sanitizer is closed; the script also runs stand-alone and checks the paths against each
other (exhaustive BM25 ranking, un-fused full-fusion)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rag_uq_b200 as rq  # noqa: E402
from rag_uq_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
n, dim, n_q = 30_011, 768, 130
engine, cdf = synth.build_synthetic_engine(n, dim, dev)
qb = synth.make_queries(n_q, n, dim, cdf, dev)
s, i = engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 10)          # seed + bound pass + window mode
full = engine.sparse.scores(qb.q_terms, qb.q_off, qb.max_terms)
assert torch.equal(s, torch.topk(full, 10, dim=1).values)
ds, di = engine.dense_topk(qb.q_emb, 50)                                          # CTA-pair kernel, bit-mask admission
torch.manual_seed(3)
router = rq.RetrievalRouter().to(dev).eval()
router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.1); router.dense_std.fill_(0.2)
router.stats_initialized = True
with torch.no_grad():
    fs, fi = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, 10, fused=True)
    us, ui = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, 10, fused=False, query_chunk=65)
    out = engine.retrieve_and_rerank(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, 10, 50, mc_samples=4, seed=1)
torch.cuda.synchronize()
assert (fi == ui).float().mean() > 0.99
print("sanitize_small ok")
