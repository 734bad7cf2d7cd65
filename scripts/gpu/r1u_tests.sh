#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? =="; tail -n 4 gpurun_out/pytest_all.log | cut -c1-200
