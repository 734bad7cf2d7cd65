#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? =="; tail -n 4 gpurun_out/pytest_all.log
show() { tail -n 1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1',d['metric'],'| value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'| bm25',round(d['kernels']['bm25_ms'],2),'dense',round(d['kernels']['dense_ms'],2),'other',round(d['kernels']['other_ms'],3),'roof',d['roofline']['kernel'][:12],round(d['roofline']['frac'],3),round(d['roofline_secondary']['frac'],3))" || tail -5 $1; }
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m.log 2>&1; echo "== 10m exit $? =="; show gpurun_out/bench_10m.log
timeout 900 python bench.py --passages 1000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1; echo "== 1m exit $? =="; show gpurun_out/bench_1m.log
timeout 600 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "== c2 exit $? =="; show gpurun_out/bench_c2.log
timeout 600 python bench.py --workload c2 --passages 10000000 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_10m.log 2>&1; echo "== c2 10m exit $? =="; show gpurun_out/bench_c2_10m.log
timeout 900 python bench.py --mode full-fusion --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ff.log 2>&1; echo "== ff exit $? =="; show gpurun_out/bench_ff.log
