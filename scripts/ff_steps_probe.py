import sys, time, json
sys.path.insert(0, '/root/repo')
import torch
import rag_uq_b200 as rq
from rag_uq_b200 import synth
dev = torch.device('cuda:0')
n = 10_000_000
engine, cdf = synth.build_synthetic_engine(n, 768, dev)
torch.manual_seed(7)
router = rq.RetrievalRouter().to(dev).eval()
router.bm25_mean.fill_(8.0); router.bm25_std.fill_(6.0); router.dense_mean.fill_(0.2); router.dense_std.fill_(0.3)
router.stats_initialized = True
batches = [synth.make_queries(1024, n, 768, cdf, dev, first_query=i * 1024) for i in range(4)]
out = []
import functools
from rag_uq_b200 import ops
timings = {}
def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    @functools.wraps(fn)
    def inner(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); timings.setdefault(label or name, []).append(round((time.perf_counter() - t) * 1e3, 2))
        return r
    setattr(obj, name, inner)
wrap(router, 'full_fusion_table'); wrap(router, 'full_fusion_envelope'); wrap(engine, '_full_fusion_exhaustive')
wrap(engine.sparse, 'scores_tiled'); wrap(ops, 'dense_mma_fused_topk'); wrap(engine, 'bm25_score_cap'); wrap(engine, '_max_passage_norm')
with torch.no_grad():
    for s in range(20):
        b = batches[s % 4]
        info, events = {}, {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.full_fusion_topk(b.q_terms, b.q_off, b.max_terms, b.q_emb, router, 10, events=events, info=info)
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ev = {k: round(v[0][0].elapsed_time(v[-1][1]), 2) for k, v in events.items()}
        out.append((s % 4, round(e0.elapsed_time(e1), 2), round(wall, 2), info.get('fallback_queries'), {k: v[-1] for k, v in timings.items()}))
for o in out: print(o)
