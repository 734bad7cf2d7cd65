#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_check.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_check.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'cpu',d['cpu_baseline'])" || tail -n 12 gpurun_out/bench_check.log
