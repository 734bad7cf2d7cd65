#!/usr/bin/env python
"""How tight is the table-term bound of bm25_kernel?  (analysis aid, B200 box)

For a batch of synthetic queries over N passages: the final admission threshold thr (the pool-th best BM25 score)
against ub_table = sum over the query's table terms of w * max impact of the row - the kernel runs its cheap
"window" phase only once thr > ub_table, the fp16 pass over EVERY document otherwise - and against the bound an
impact-capped table would give (rows capped at their q-quantile impact, the few documents above the cap listed
separately as markers).  Prints one JSON line.

    python scripts/analyze_bm25_bounds.py --passages 10000000 --batch 1024 --pool 50
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--passages", type=int, default=10_000_000)
    p.add_argument("--batch", type=int, default=1024)
    p.add_argument("--pool", type=int, default=50)
    a = p.parse_args()
    from rag_uq_b200 import synth
    dev = torch.device("cuda:0")
    engine, cdf = synth.build_synthetic_engine(a.passages, 768, dev, with_dense=False)
    sh = engine.sparse
    qb = synth.make_queries(a.batch, a.passages, 768, cdf, dev)
    score, ids = sh.score_topk(qb.q_terms, qb.q_off, qb.max_terms, a.pool)
    seed = sh.seed(qb.q_terms, qb.q_off, qb.max_terms, a.pool)
    thr = torch.where(ids[:, -1] >= 0, score[:, -1], torch.zeros_like(score[:, -1]))
    terms = qb.q_terms.view(a.batch, -1).long()
    ok = (terms >= 0) & (terms < sh.vocab)
    w = torch.where(ok, sh.idf[terms.clamp(0, sh.vocab - 1)] * (sh.k1 + 1.0), torch.zeros((), device=dev))
    # table row of every query term (-1 = posting-list term)
    row = torch.full_like(terms, -1)
    pos = torch.searchsorted(sh.dense_terms.long(), terms.clamp(0, sh.vocab - 1))
    hit = ok & (pos < sh.dense_terms.numel()) & (sh.dense_terms.long()[pos.clamp(max=sh.dense_terms.numel() - 1)] == terms)
    row[hit] = pos[hit]
    imp = sh.dense_imp[:, :sh.n_docs]
    out = {"passages": a.passages, "batch": a.batch, "pool": a.pool, "table_rows": int(sh.dense_terms.numel()),
           "table_terms_per_query": float(hit.float().sum(1).mean()), "list_terms_per_query": float((ok & ~hit).float().sum(1).mean())}

    def bound(cap_per_row):
        c = torch.where(hit, cap_per_row[row.clamp(min=0)], torch.zeros((), device=dev))
        return (w.clamp(min=0) * c).sum(1)

    ub = bound(sh.dense_maximp)
    out["thr_mean"], out["ub_table_mean"], out["seed_mean"] = float(thr.mean()), float(ub.mean()), float(seed.mean())
    out["frac_thr_above_ub"] = float((thr > ub).float().mean())
    out["frac_seed_above_ub"] = float((seed > ub).float().mean())
    rows_used = torch.unique(row[hit])
    for q in (0.99, 0.999, 0.9999):
        cap = sh.dense_maximp.clone()
        k = max(1, int(round((1.0 - q) * sh.n_docs)))
        for r in rows_used.tolist():
            cap[r] = torch.topk(imp[r].float(), k).values[-1]
        ubq = bound(cap)
        out[f"cap_q{q}"] = {"ub_mean": float(ubq.mean()), "frac_thr_above_ub": float((thr > ubq).float().mean()),
                            "marker_postings_per_query": float(hit.float().sum(1).mean() * k)}
    # list-term postings per query for scale
    toff = sh.term_off
    lt = torch.where(ok & ~hit, toff[terms.clamp(0, sh.vocab - 1) + 1] - toff[terms.clamp(0, sh.vocab - 1)], torch.zeros((), dtype=torch.int64, device=dev))
    out["list_postings_per_query"] = float(lt.sum(1).float().mean())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
