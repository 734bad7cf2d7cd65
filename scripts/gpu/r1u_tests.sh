#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -k "window_mode or bm25" > gpurun_out/pytest_win.log 2>&1
echo "== pytest exit $? =="; tail -n 25 gpurun_out/pytest_win.log
