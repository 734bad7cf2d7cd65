#!/bin/bash
mkdir -p gpurun_out
for d in 0 4; do
RAGB_FF_DEBUG=$d timeout 600 python scripts/bench_full_fusion.py 10000000 1024 10 > gpurun_out/ff_dbg$d.log 2>&1; echo "== dbg $d exit $? =="; tail -n 1 gpurun_out/ff_dbg$d.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('fused_gemm_ms','bm25_scores_ms','gate_evaluations','full_fusion_ms')})"
done
