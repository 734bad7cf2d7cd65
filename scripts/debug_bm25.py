"""Debug aid: BM25 pruning statistics for one batch (RAGB_BM25_DEBUG=1)."""
import ctypes, os, sys
from pathlib import Path
os.environ["RAGB_BM25_DEBUG"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import rag_uq_b200 as rq
from rag_uq_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
engine, cdf = synth.build_synthetic_engine(n, 64, dev, with_dense=False)
qb = synth.make_queries(1024, n, 64, cdf, dev)
lib = rq._lib.lib
out = (ctypes.c_ulonglong * 8)()
lib.ragb_debug_bm25_counters(out)
s, i = engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50)
torch.cuda.synchronize()
lib.ragb_debug_bm25_counters(out)
full, pr0, pr1, docs, seeded, strong, approx_ok, approx_fail = [int(x) for x in out[:8]]
print(f"table rows {engine.sparse.dense_terms.numel()}  super-ranges: full {full}  pruned-empty {pr0}  pruned-with-postings {pr1}  "
      f"docs scored in pruned mode {docs} ({docs / max(pr1, 1):.1f} per super-range)  queries seeded {seeded}  seed>bound {strong}  "
      f"fp16 bound pass: decided {approx_ok}, fell back {approx_fail}")
for _ in range(3):
    engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50)
e1.record()
torch.cuda.synchronize()
print(f"bm25 pool-50 x 1024 queries (debug counters on): {e0.elapsed_time(e1) / 5:.2f} ms")
print("kth scores (first 8 queries):", s[:8, -1].tolist())
idf = engine.sparse.idf
terms = qb.q_terms.view(1024, -1)[:8].long().clamp(0, engine.sparse.vocab - 1)
print("weights of first query:", (idf[terms[0]] * 2.5).tolist(), "df:", (engine.sparse.term_off[terms[0] + 1] - engine.sparse.term_off[terms[0]]).tolist())
