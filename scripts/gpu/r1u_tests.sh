#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider -k "random_small_corpora" > gpurun_out/pytest_rand.log 2>&1
echo "== pytest exit $? =="; tail -n 30 gpurun_out/pytest_rand.log | cut -c1-220
