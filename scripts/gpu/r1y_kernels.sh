#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_kernels.py 10000000 > gpurun_out/kernels_10m.json 2> gpurun_out/kernels_10m.err; echo "exit $?"; tail -c 2500 gpurun_out/kernels_10m.json
