#!/usr/bin/env python
"""recall@k of the reference's approximate dense path (an HNSW with ChromaDB's default parameters, restated in
oracle/hnsw.py because chromadb / hnswlib are not installable here) against this repository's exact search on the
synthetic corpus.  The exact ids come from the tcgen05 kernel; the HNSW build is CPU work (minutes beyond ~20k rows).

    python tests/hnsw_recall_report.py [passages ...]
"""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rag_uq_b200 as rq  # noqa: E402
from oracle import hnsw  # noqa: E402  (lives under tests/: the oracle may only be used from there)
from rag_uq_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
for n in [int(a) for a in sys.argv[1:]] or [10_000]:
    passages = synth.passage_embeddings(0, n, 768, dev)
    qb = synth.make_queries(128, n, 768, synth.zipf_cdf(synth.vocab_size(n), dev), dev)
    _, ids = rq.ops.dense_mma_topk(passages, qb.q_emb, 10, 0, 3)
    t0 = time.time()
    rep = hnsw.recall_report(passages.float().cpu().numpy(), qb.q_emb.float().cpu().numpy(), ids.cpu().numpy(), 10)
    rep["hnsw_build_and_search_seconds"] = round(time.time() - t0, 1)
    print(json.dumps(rep), flush=True)
