#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "dense or end_to_end or row_sharded" > gpurun_out/pytest_dense.log 2>&1
echo "== pytest dense exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 5 gpurun_out/pytest_dense.log
for v in 0 1; do
  timeout 600 python bench.py --passages 1000000 --steps 10 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_1m_v$v.log 2>&1
  echo "== bench1m v$v exit $? ==" | tee -a gpurun_out/summary.txt; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_1m_v$v.log').read().strip().splitlines()[-1])
print('value',round(d['value']),'dense_ms',round(d['kernels']['dense_ms'],2),'bm25_ms',round(d['kernels']['bm25_ms'],2),'frac',round(d['roofline']['frac'],3))"
  timeout 900 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_10m_v$v.log 2>&1
  echo "== bench10m v$v exit $? ==" | tee -a gpurun_out/summary.txt; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_10m_v$v.log').read().strip().splitlines()[-1])
print('value',round(d['value']),'dense_ms',round(d['kernels']['dense_ms'],2),'bm25_ms',round(d['kernels']['bm25_ms'],2),'frac',round(d['roofline']['frac'],3))"
done
cat gpurun_out/summary.txt
