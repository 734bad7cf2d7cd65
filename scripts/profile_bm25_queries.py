"""Where does bm25_kernel spend its warp cycles, per query and per phase?  (diagnostic, B200 box)

Needs the instrumented build (-DRAGB_BM25_PROFILE; build it HERE first, it travels with the snapshot):
  python efficient-rag-with-learned-retrieval-and-uncertainty-quantification_b200/build.py --profile
  gpurun -- 'python scripts/profile_bm25_queries.py [passages]'        (loads libragb200_prof.so via RAGB_LIB_NAME)
Prints the share of each phase and how the cost is distributed over the 1024 queries of one batch.
"""
import ctypes
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("RAGB_LIB_NAME", "libragb200_prof.so")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import rag_uq_b200 as rq
from rag_uq_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
engine, cdf = synth.build_synthetic_engine(n, 64, dev, with_dense=False)
qb = synth.make_queries(1024, n, 64, cdf, dev)
lib = rq._lib.lib
buf = (ctypes.c_ulonglong * (8 * 4096))()
for _ in range(2):
    engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50)
torch.cuda.synchronize()
assert lib.ragb_debug_bm25_profile(buf) == 0
s, _ = engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50)
torch.cuda.synchronize()
assert lib.ragb_debug_bm25_profile(buf) == 0
c = np.frombuffer(buf, dtype=np.uint64).reshape(8, 4096)[:, :1024].astype(np.float64)
names = ["setup", "window", "dense_or_bound", "fold", "win_stream", "win_compact", "win_score", "fold_wait_for_slowest_warp"]
total = c[:4].sum()
out = {"passages": n, "warp_cycles_total": total, "share": {names[i]: c[i].sum() / total for i in range(8)}}
per_q = c[:4].sum(axis=0)
order = np.argsort(-per_q)
cum = np.cumsum(per_q[order]) / total
out["queries_for_share"] = {f"{int(p * 100)}%": int(np.searchsorted(cum, p) + 1) for p in (0.25, 0.5, 0.75, 0.9)}
out["per_query_cycles"] = {"max": per_q.max(), "median": float(np.median(per_q)), "mean": per_q.mean()}
# the costliest queries: which phase, which terms (df)
idf = engine.sparse.idf
toff = engine.sparse.term_off
terms = qb.q_terms.view(1024, -1).long().clamp(0, engine.sparse.vocab - 1)
top = []
for qi in order[:12]:
    df = (toff[terms[qi] + 1] - toff[terms[qi]]).tolist()
    top.append({"q": int(qi), "share_of_total": per_q[qi] / total,
                "phases": {names[i]: c[i, qi] / per_q[qi] for i in range(4)}, "df": df, "kth": float(s[qi, -1])})
out["costliest"] = top
# phase split of the heavy (dense_or_bound-dominated) queries vs the rest
heavy = c[2] > 0.5 * per_q
out["bound_pass_dominated"] = {"queries": int(heavy.sum()), "share_of_total": per_q[heavy].sum() / total}
print(json.dumps(out, indent=1, default=float))
