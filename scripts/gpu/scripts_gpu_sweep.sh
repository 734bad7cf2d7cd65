#!/bin/bash
mkdir -p gpurun_out
for f in 4 6 8 12 16 32 64; do
  RAGB_DENSE_MIN_FRACTION=$f timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/sweep_$f.log 2>&1
  python -c "
import json
d=json.loads(open('gpurun_out/sweep_$f.log').read().strip().splitlines()[-1])
print('fraction 1/$f: bm25_ms', round(d['kernels']['bm25_ms'],2), 'value', round(d['value']))" || tail -3 gpurun_out/sweep_$f.log
done
