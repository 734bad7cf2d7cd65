#!/bin/bash
# full GPU suite + headline bench (with CPU baseline) + reference arm + HNSW recall report
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? =="; tail -n 4 gpurun_out/pytest_all.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "== bench default exit $? =="; tail -n 1 gpurun_out/bench_default.log | cut -c1-2500
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "== bench reference exit $? =="; tail -n 1 gpurun_out/bench_reference.log | cut -c1-900
timeout 900 python tests/hnsw_recall_report.py 10000 20000 > gpurun_out/hnsw_recall.log 2>&1; echo "== hnsw exit $? =="; tail -n 2 gpurun_out/hnsw_recall.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
