#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --passages 1000000 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bm25_kernel' -s 3 -c 1 -o gpurun_out/prof_bm25 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
