#!/usr/bin/env python
"""Stand-alone timing of the tcgen05 dense kernel (CUDA events) over shard sizes, batch sizes, variants and k.
Needs a B200.   python scripts/bench_mma.py [rows ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rag_uq_b200 as rq  # noqa: E402,F401
from rag_uq_b200 import ops, synth  # noqa: E402

import os
sizes = [int(a) for a in sys.argv[1:]] or [10_000_000]
VARIANTS = [int(v) for v in os.environ.get('MMA_VARIANTS', '2,3').split(',')]
KS = [int(v) for v in os.environ.get('MMA_KS', '1,10,50').split(',')]
dev = torch.device("cuda:0")
for n in sizes:
    passages = synth.passage_embeddings(0, n, 768, dev)
    cdf = synth.zipf_cdf(synth.vocab_size(n), dev)
    for b in (1024,):
        qb = synth.make_queries(b, n, 768, cdf, dev)
        for variant in VARIANTS:
            for k in KS:
                for _ in range(3):
                    ops.dense_mma_topk(passages, qb.q_emb, k, 0, variant)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    ops.dense_mma_topk(passages, qb.q_emb, k, 0, variant)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"N={n} B={b} variant={variant} k={k}: {ms:.3f} ms  {2.0 * b * n * 768 / ms / 1e9:.0f} TFLOP/s", flush=True)
    del passages
