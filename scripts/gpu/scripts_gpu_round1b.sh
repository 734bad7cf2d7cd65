#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 15 gpurun_out/pytest_all.log
for v in 0 1; do
  timeout 600 python bench.py --passages 1000000 --steps 10 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_1m_v$v.log 2>&1
  echo "== bench1m v$v exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_1m_v$v.log | cut -c1-1800
done
for v in 0 1; do
  timeout 900 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_10m_v$v.log 2>&1
  echo "== bench10m v$v exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_10m_v$v.log | cut -c1-1800
done
cat gpurun_out/summary.txt
