#!/usr/bin/env python
"""BM25 pool-k kernel, same process A/B: with and without the fp16 impact bounds of the table terms (CUDA events)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rag_uq_b200 as rq  # noqa: E402,F401
from rag_uq_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
engine, cdf = synth.build_synthetic_engine(n, 64, dev, with_dense=False)
sp = engine.sparse
batches = [synth.make_queries(1024, n, 64, cdf, dev, first_query=i * 1024) for i in range(4)]


def timed(reps=8):
    for b in batches[:2]:
        sp.score_topk(b.q_terms, b.q_off, b.max_terms, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        b = batches[r % 4]
        sp.score_topk(b.q_terms, b.q_off, b.max_terms, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


keep = sp.dense_imp, sp.dense_maximp
for rnd in range(2):
    sp.dense_imp, sp.dense_maximp = keep
    t_with = timed()
    sp.dense_imp, sp.dense_maximp = keep[0][:0], keep[1][:0]
    t_without = timed()
    print(f"N={n} k={k} round {rnd}: with impact bounds {t_with:.2f} ms, without {t_without:.2f} ms", flush=True)
