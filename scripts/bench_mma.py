#!/usr/bin/env python
"""Stand-alone timing of the tcgen05 dense kernel variants (CUDA events).  Needs a B200."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rag_uq_b200 as rq  # noqa: E402
from rag_uq_b200 import ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
passages = synth.passage_embeddings(0, n, 768, dev)
cdf = synth.zipf_cdf(synth.vocab_size(n), dev)
for b in (256, 1024):
    qb = synth.make_queries(b, n, 768, cdf, dev)
    for variant in (0, 2, 3):
        for k in (10, 50):
            for _ in range(3):
                ops.dense_mma_topk(passages, qb.q_emb, k, 0, variant)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.dense_mma_topk(passages, qb.q_emb, k, 0, variant)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"B={b} variant={variant} k={k}: {ms:.2f} ms  {2.0 * b * n * 768 / ms / 1e9:.0f} TFLOP/s", flush=True)
