#!/bin/bash
# end-of-round verification at HEAD: full GPU suite, smoke, default bench (with CPU baseline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? =="; tail -n 4 gpurun_out/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "== bench default exit $? =="; tail -n 1 gpurun_out/bench_default.log | cut -c1-1800
