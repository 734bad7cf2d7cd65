#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "dense or end_to_end or sharded or large_corpus" > gpurun_out/pytest_dense.log 2>&1
echo "== pytest dense exit $? =="; tail -n 3 gpurun_out/pytest_dense.log
MMA_VARIANTS=3 timeout 900 python scripts/bench_mma.py 1250000 10000000 > gpurun_out/mma_sizes.log 2>&1; echo "exit $?"; cat gpurun_out/mma_sizes.log
