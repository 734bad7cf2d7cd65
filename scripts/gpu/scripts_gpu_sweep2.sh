#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -k "bm25 or end_to_end or sharded or dropin or incremental or full_fusion" > gpurun_out/pytest_bm25.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_bm25.log
for f in 8 16 32 64 128; do
  RAGB_DENSE_MIN_FRACTION=$f timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/sweep_$f.log 2>&1
  python -c "
import json
d=json.loads(open('gpurun_out/sweep_$f.log').read().strip().splitlines()[-1])
print('fraction 1/$f: bm25_ms', round(d['kernels']['bm25_ms'],2), 'dense_ms', round(d['kernels']['dense_ms'],2), 'value', round(d['value']))" || tail -3 gpurun_out/sweep_$f.log
done
