#!/bin/bash
mkdir -p gpurun_out
for w in 128 192 256 320 400; do echo "window target=$w"; RAGB_BM25_WINDOW=$w timeout 600 python scripts/bench_bm25.py 10000000 50 2>&1 | tail -1; done
