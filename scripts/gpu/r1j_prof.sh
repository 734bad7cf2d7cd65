#!/bin/bash
# round-1 profile of the headline bench step: launch list + one --set full capture of the three hot kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
tail -n 1 gpurun_out/prof_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_10m_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'bm25_kernel|bm25_seed_kernel|dense_mma_pair_kernel' -s 9 -c 3 -o gpurun_out/prof_r01d_10m -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/prof_r01d_10m.ncu-rep --page raw --csv > gpurun_out/r01d_ncu_full_10m_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01d_10m.ncu-rep --page source --csv --kernel-name regex:'bm25_kernel' > gpurun_out/r01d_src_bm25.csv 2>/dev/null
ncu -i gpurun_out/prof_r01d_10m.ncu-rep --page source --csv --kernel-name regex:'dense_mma_pair_kernel' > gpurun_out/r01d_src_mma.csv 2>/dev/null
ls -la gpurun_out | grep r01d
