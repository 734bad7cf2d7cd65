#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "c2 exit $?"; tail -1 gpurun_out/bench_c2.log | cut -c1-2500
timeout 600 python bench.py --workload c2 --passages 10000000 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_10m.log 2>&1; echo "c2 10m exit $?"; tail -1 gpurun_out/bench_c2_10m.log | cut -c1-2500
