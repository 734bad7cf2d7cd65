#!/bin/bash
# Final round-1 evidence at the bench configuration (10M passages): launch list + full capture of both hot kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof10m_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bm25|dense_mma|gemv|topk|hybrid_fuse|router|stats|idf|norm_kernel|sort_rows' -c 300 --csv --log-file gpurun_out/launches_10m.csv $CMD > gpurun_out/ncu_launches_10m.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof10m_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_mma_kernel|bm25_kernel' -s 6 -c 2 -o gpurun_out/prof_hot_10m -f $CMD > gpurun_out/ncu_full_10m.log 2>&1
echo "full capture exit $?"
tail -2 gpurun_out/ncu_full_10m.log
