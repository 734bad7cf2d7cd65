#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'bm25_kernel' -s 3 -c 1 -o gpurun_out/prof_bm25_win -f python scripts/bench_bm25.py 10000000 50 > gpurun_out/ncu_bm25_win.log 2>&1
echo "exit $?"
ncu -i gpurun_out/prof_bm25_win.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/bm25_win_mix.csv 2>/dev/null
ncu -i gpurun_out/prof_bm25_win.ncu-rep --page raw --csv > gpurun_out/bm25_win_raw.csv 2>/dev/null
ls -la gpurun_out/bm25_win_mix.csv
