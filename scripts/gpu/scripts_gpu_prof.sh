#!/bin/bash
# ncu: launch list + one full capture of the two hot kernels, on the 1M-passage configuration.
mkdir -p gpurun_out
CMD="python bench.py --passages 1000000 --steps 2 --warmup 3 --variant ${VARIANT:-1} --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ragb|dense_mma|bm25|topk|hybrid_fuse|router|gemv' -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_mma_kernel|bm25_kernel' -s 4 -c 2 -o gpurun_out/prof_hot -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/
