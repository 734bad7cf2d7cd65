#!/bin/bash
# fused full-fusion epilogue: parity tests, then timing at 1M and 10M
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "full_fusion" > gpurun_out/pytest_ff.log 2>&1
echo "== pytest ff exit $? =="; tail -n 15 gpurun_out/pytest_ff.log
timeout 600 python scripts/bench_full_fusion.py 1000000 1024 10 > gpurun_out/ff_1m.log 2>&1; echo "== ff 1m exit $? =="; tail -n 3 gpurun_out/ff_1m.log
timeout 900 python scripts/bench_full_fusion.py 10000000 1024 10 > gpurun_out/ff_10m.log 2>&1; echo "== ff 10m exit $? =="; tail -n 3 gpurun_out/ff_10m.log
