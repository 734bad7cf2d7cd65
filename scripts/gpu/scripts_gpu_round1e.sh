#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 5 gpurun_out/pytest_all.log
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1 value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),'dense_ms',round(d['kernels']['dense_ms'],2),'bm25_ms',round(d['kernels']['bm25_ms'],2),'other',round(d['kernels']['other_ms'],2),'clocks',d['clocks'])" || tail -5 $1; }
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-overlap > gpurun_out/bench_10m_seq.log 2>&1; echo "== seq exit $? ==" | tee -a gpurun_out/summary.txt; show gpurun_out/bench_10m_seq.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m_ovl.log 2>&1; echo "== overlap exit $? ==" | tee -a gpurun_out/summary.txt; show gpurun_out/bench_10m_ovl.log
RAGB_MMA_STAGES=4 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-overlap > gpurun_out/bench_10m_seq4.log 2>&1; echo "== seq 4 stages exit $? ==" | tee -a gpurun_out/summary.txt; show gpurun_out/bench_10m_seq4.log
timeout 900 python bench.py --passages 1000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m_ovl.log 2>&1; echo "== 1m overlap exit $? ==" | tee -a gpurun_out/summary.txt; show gpurun_out/bench_1m_ovl.log
cat gpurun_out/summary.txt
