#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "dense_mma" -x > gpurun_out/pytest_mma.log 2>&1; echo "pytest mma exit $?"; tail -15 gpurun_out/pytest_mma.log
for v in 2 3; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/bench_10m_v$v.log 2>&1; echo "bench v$v exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_10m_v$v.log').read().strip().splitlines()[-1])
print('variant $v value',round(d['value']),'dense_ms',round(d['kernels']['dense_ms'],2),'bm25_ms',round(d['kernels']['bm25_ms'],2))" || tail -5 gpurun_out/bench_10m_v$v.log
done
