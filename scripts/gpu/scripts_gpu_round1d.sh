#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 25 gpurun_out/pytest_all.log
for n in 1000000 10000000; do
  timeout 900 python bench.py --passages $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$n.log 2>&1
  echo "== bench $n exit $? ==" | tee -a gpurun_out/summary.txt; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$n.log').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'dense_ms',round(d['kernels']['dense_ms'],2),'bm25_ms',round(d['kernels']['bm25_ms'],2),'frac',round(d['roofline']['frac'],3), 'build_s', d['config']['build_seconds'])" || tail -5 gpurun_out/bench_$n.log
done
cat gpurun_out/summary.txt
