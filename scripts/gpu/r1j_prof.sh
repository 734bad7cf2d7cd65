#!/bin/bash
# end-of-round profile of the headline bench step: NVTX-filtered launch list + one --set full capture of the hot kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
tail -n 1 gpurun_out/prof_plain.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bench_timed/" -c 400 --csv --log-file gpurun_out/r01g_launches_10m_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'bm25_kernel|bm25_seed_kernel|dense_mma_pair_kernel' -s 9 -c 3 -o gpurun_out/prof_r01g_10m -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/prof_r01g_10m.ncu-rep --page raw --csv > gpurun_out/r01g_ncu_full_10m_raw.csv 2>/dev/null
ls -la gpurun_out | grep r01f
