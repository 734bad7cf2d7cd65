#!/bin/bash
mkdir -p gpurun_out
show() { tail -n 1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'| bm25',round(d['kernels']['bm25_ms'],2),'dense',round(d['kernels']['dense_ms'],2),'other',round(d['kernels']['other_ms'],3))" || tail -5 $1; }
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_seq.log 2>&1; show gpurun_out/bench_seq.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --overlap > gpurun_out/bench_ovl.log 2>&1; show gpurun_out/bench_ovl.log
RAGB_MMA_STAGES=5 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --overlap > gpurun_out/bench_ovl5.log 2>&1; show gpurun_out/bench_ovl5.log
