#!/usr/bin/env python
"""Per-kernel timings (CUDA events, L2-exceeding inputs) for profiles/README.md.  Needs a B200."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rag_uq_b200 as rq  # noqa: E402
from rag_uq_b200 import ops, synth  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    dev = torch.device("cuda:0")
    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) \
        if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0}
    engine, cdf = synth.build_synthetic_engine(n, 768, dev)
    out = {"passages": n}
    for b in (1, 4, 8):
        qb = synth.make_queries(b, n, 768, cdf, dev)
        ms = timeit(lambda: ops.dense_gemv_topk(engine.passages, qb.q_emb, 50, 0))
        gb = n * 768 * 2 * (-(-b // 4)) / 1e9
        out[f"gemv_b{b}_k50"] = {"ms": ms, "GB/s": gb / ms * 1e3, "frac_hbm": gb / ms * 1e3 / peaks["hbm_gbs"]}
    for b in (128, 256, 1024):
        qb = synth.make_queries(b, n, 768, cdf, dev)
        for variant in (0, 1, 2):
            for k in (10, 50):
                ms = timeit(lambda: ops.dense_mma_topk(engine.passages, qb.q_emb, k, 0, variant), iters=5)
                tf = 2.0 * b * n * 768 / ms / 1e9
                out[f"mma_v{variant}_b{b}_k{k}"] = {"ms": ms, "TFLOP/s": tf, "frac_sustained": tf / peaks["bf16_tflops_sustained"],
                                                     "frac_burst": tf / peaks["bf16_tflops"],
                                                     "GB/s_hbm": n * 768 * 2 / ms / 1e6}
    for b in (1, 64, 1024):
        qb = synth.make_queries(b, n, 768, cdf, dev)
        ms = timeit(lambda: engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50), iters=5)
        qt = qb.q_terms.long()
        ok = (qt >= 0) & (qt < engine.sparse.vocab)
        df = int((engine.sparse.term_off[qt[ok] + 1] - engine.sparse.term_off[qt[ok]]).sum())
        out[f"bm25_b{b}_k50"] = {"ms": ms, "postings": df, "Gpostings/s": df / ms / 1e6, "algorithmic_GB/s": df * 6 / ms / 1e6,
                                 "frac_hbm": df * 6 / ms / 1e6 / peaks["hbm_gbs"], "dense_table_rows": int(engine.sparse.dense_terms.numel())}
    # candidate-list kernels on the C4 shape: 1024 queries x 100 candidates, T = 30
    g = torch.Generator(device="cpu").manual_seed(0)
    b100 = (torch.rand(1024, 100, generator=g) * 20).to(dev)
    d100 = torch.rand(1024, 100, generator=g).to(dev)
    router = rq.RetrievalRouter().to(dev).eval()
    with torch.no_grad():
        out["router_forward_1024x100"] = {"ms": timeit(lambda: router(b100, d100))}
        out["router_rerank_1024x100_k10"] = {"ms": timeit(lambda: router.hybrid_rerank(b100, d100, 10))}
        out["mc_dropout_T30_1024x100"] = {"ms": timeit(lambda: router.mc_dropout(b100, d100, n_samples=30, seed=1))}
        out["mc_dropout_T30_1x100"] = {"ms": timeit(lambda: router.mc_dropout(b100[:1].contiguous(), d100[:1].contiguous(), n_samples=30, seed=1))}
    bs, bi = engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50)
    ds, di = ops.dense_mma_topk(engine.passages, qb.q_emb, 50, 0, 0)
    out["hybrid_fuse_1024_pool50"] = {"ms": timeit(lambda: ops.hybrid_fuse_topk(bs, bi, ds, di, 10))}
    stack_s, stack_i = torch.stack([bs] * 8, 1).contiguous(), torch.stack([bi] * 8, 1).contiguous()
    out["topk_merge_1024x8x50"] = {"ms": timeit(lambda: ops.topk_merge(stack_s, stack_i, 50))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
