#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider -k "fused_epilogue_random_shapes" > gpurun_out/pytest_rand.log 2>&1
echo "== pytest exit $? =="; tail -n 40 gpurun_out/pytest_rand.log | cut -c1-200
