#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 5 gpurun_out/pytest_all.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1 value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'dense_ms',round(d['kernels']['dense_ms'],3),'bm25_ms',round(d['kernels']['bm25_ms'],3),'other',round(d['kernels']['other_ms'],3),'clocks',d['clocks'])" || tail -5 $1; }
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m.log 2>&1; echo "== 10m exit $? ==" | tee -a gpurun_out/summary.txt; show gpurun_out/bench_10m.log
timeout 900 python bench.py --passages 1000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1; echo "== 1m exit $? ==" | tee -a gpurun_out/summary.txt; show gpurun_out/bench_1m.log
timeout 600 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "== c2 exit $? ==" | tee -a gpurun_out/summary.txt; show gpurun_out/bench_c2.log
cat gpurun_out/summary.txt
