#!/bin/bash
# 1-GPU validation of the bench modes before the 8-GPU run + NVTX-filtered launch list of the timed steps
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $? =="; tail -n 1 gpurun_out/$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['metric'],'| value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'| bm25',round(d['kernels']['bm25_ms'],2),'dense',round(d['kernels']['dense_ms'],2),'launches',d['gpu_launches'],'roof',d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['traffic'])" || tail -n 5 gpurun_out/$name.log; }
run bench_c5small --workload c5 --passages 4000000 --steps 3 --warmup 3 --no-cpu-baseline
run bench_c4 --workload c4 --steps 3 --warmup 3 --no-cpu-baseline
run bench_ff --mode full-fusion --steps 3 --warmup 3 --no-cpu-baseline
run bench_c3 --steps 5 --warmup 3
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bench_timed/" -c 400 --csv --log-file gpurun_out/r01d_launches_10m_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; grep -c . gpurun_out/r01d_launches_10m_bench.csv
