#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x -k "bm25 or end_to_end or sharded or large_corpus" > gpurun_out/pytest_bm25.log 2>&1
echo "== pytest bm25 exit $? =="; tail -n 5 gpurun_out/pytest_bm25.log
timeout 600 python scripts/bench_bm25.py 10000000 50 2>&1 | tail -1
timeout 600 python scripts/bench_bm25.py 1250000 50 2>&1 | tail -1
timeout 600 python scripts/debug_bm25.py 10000000 2>&1 | head -1 | cut -c1-500
