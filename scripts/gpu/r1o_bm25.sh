#!/bin/bash
mkdir -p gpurun_out
for b in 72 96 144 192 288; do echo "blocks per sm=$b"; RAGB_BM25_BLOCKS_PER_SM=$b timeout 600 python scripts/bench_bm25.py 10000000 50 2>&1 | tail -1; done
for b in 12 48 96 192; do echo "blocks per sm=$b"; RAGB_BM25_BLOCKS_PER_SM=$b timeout 600 python scripts/bench_bm25.py 1250000 50 2>&1 | tail -1; done
