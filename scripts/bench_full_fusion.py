#!/usr/bin/env python
"""Full-fusion mode at scale (CUDA events): BM25 get_scores matrix + tcgen05 GEMM with gate / fusion / top-k in
the epilogue, against the dense-only kernel on the same shapes.  Needs a B200.

    python scripts/bench_full_fusion.py [passages] [queries] [k]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rag_uq_b200 as rq  # noqa: E402
from rag_uq_b200 import ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
engine, cdf = synth.build_synthetic_engine(n, 768, dev)
torch.manual_seed(7)
router = rq.RetrievalRouter().to(dev).eval()
router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.2); router.dense_std.fill_(0.3)
router.stats_initialized = True
qb = synth.make_queries(n_q, n, 768, cdf, dev)
counters = torch.zeros(2, dtype=torch.int64, device=dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"passages": n, "queries": n_q, "k": k}
with torch.no_grad():
    out["dense_only_ms"] = timed(lambda: ops.dense_mma_topk(engine.passages, qb.q_emb, k, 0, 3))
    events = {}
    counters.zero_()
    engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=True, counters=counters)
    torch.cuda.synchronize()
    out["gate_evaluations"], out["admissions"] = (int(v) for v in counters.tolist())
    out["pairs"] = n * n_q
    out["full_fusion_ms"] = timed(lambda: engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k,
                                                                   fused=True), reps=3, warm=1)
    engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=True, events=events)
    torch.cuda.synchronize()
    out["bm25_scores_ms"] = sum(a.elapsed_time(b) for a, b in events["bm25"])
    out["fused_gemm_ms"] = sum(a.elapsed_time(b) for a, b in events["dense"])
    out["fused_gemm_tflops"] = 2.0 * n * n_q * 768 / out["fused_gemm_ms"] / 1e9
    out["queries_per_s"] = n_q / out["full_fusion_ms"] * 1e3
    if n * n_q <= 4_000_000_000:   # cross-check against the un-fused path where its matrices fit comfortably
        fs, fi = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=True)
        us, ui = engine.full_fusion_topk(qb.q_terms, qb.q_off, qb.max_terms, qb.q_emb, router, k, fused=False, query_chunk=32)
        out["ids_equal_frac"] = float((fi == ui).float().mean())
        out["max_score_diff"] = float((fs - us).abs().max())
print(json.dumps(out), flush=True)
