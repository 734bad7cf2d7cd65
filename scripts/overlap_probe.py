#!/usr/bin/env python
"""Can blocks of another kernel become co-resident with the tcgen05 GEMM blocks?  (diagnostic, B200 box)

Runs the dense kernel (variant 4: CTA pairs, 4-stage ring, 129 KB of shared memory, 120 registers x 320 threads per SM)
on a high-priority stream and, started right after its sampled prefix, a co-runner on the default stream:
  * torch elementwise kernels (tiny footprint: no shared memory, ~32 registers),
  * the BM25 top-k search (8 warps, 80 registers, 53 KB of shared memory per block).
Prints alone / together times: together ~ max(alone) means the blocks share the SMs, together ~ sum means they do not.

What it found: with the GEMM at 120 registers the BM25 blocks never became co-resident - registers
are allocated per warp in units of 256 and a block's warps are rounded up to a multiple of four, so the GEMM block held
12 x 3 840 of the SM's 65 536 registers and the 8 x 2 560 of a BM25 block were 1 024 too many.  Since the GEMM needs 109
(-> 112) registers one BM25 block per SM fits: 4M passages, dense 5.12 ms + BM25 5.84 ms alone, 9.92 ms together.
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from rag_uq_b200 import ops, synth
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    engine, cdf = synth.build_synthetic_engine(n, 768, dev)
    qb = synth.make_queries(1024, n, 768, cdf, dev)
    side = torch.cuda.Stream(device=dev, priority=-1)
    cur = torch.cuda.current_stream()
    big = torch.ones(1 << 28, dtype=torch.float32, device=dev)     # 1 GiB

    def dense(variant=4):
        thr, ws = ops.dense_mma_sample(engine.passages, qb.q_emb, 50, 0, variant)
        ev = torch.cuda.Event()
        ev.record()
        out = ops.dense_mma_seeded(engine.passages, qb.q_emb, 50, 0, variant, thr, ws)
        return ev, out

    def co_torch(reps=12):
        for _ in range(reps):
            big.mul_(1.0000001)

    def co_bm25():
        engine.sparse.score_topk(qb.q_terms, qb.q_off, qb.max_terms, 50)

    def timed(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    def together(co):
        def run():
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ev, _ = dense()
            cur.wait_event(ev)
            co()
            cur.wait_stream(side)
        return run

    out = {"passages": n}
    for _ in range(2):
        out["dense_v4_alone_ms"] = timed(lambda: dense(4))
        out["dense_v3_alone_ms"] = timed(lambda: dense(3))
        out["torch_alone_ms"] = timed(co_torch)
        out["bm25_alone_ms"] = timed(co_bm25)
        out["dense_plus_torch_ms"] = timed(together(co_torch))
        out["dense_plus_bm25_ms"] = timed(together(co_bm25))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
