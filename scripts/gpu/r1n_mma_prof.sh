#!/bin/bash
mkdir -p gpurun_out
export MMA_VARIANTS=3 MMA_KS=50
ncu --set full --clock-control none --import-source on -k regex:'dense_mma_pair_kernel' -s 5 -c 1 -o gpurun_out/prof_mma_small -f python scripts/bench_mma.py 1250000 > gpurun_out/ncu_mma_small.log 2>&1
echo "exit $?"
ncu -i gpurun_out/prof_mma_small.ncu-rep --page source --csv > gpurun_out/src_mma_small.csv 2>/dev/null
ncu -i gpurun_out/prof_mma_small.ncu-rep --page raw --csv > gpurun_out/raw_mma_small.csv 2>/dev/null
ls -la gpurun_out/src_mma_small.csv
