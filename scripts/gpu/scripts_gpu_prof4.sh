#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof10m_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bm25_kernel' -s 3 -c 1 -o gpurun_out/prof_bm25_10m -f $CMD > gpurun_out/ncu_full_10m.log 2>&1
echo "full capture exit $?"
