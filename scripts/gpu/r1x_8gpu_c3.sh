#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu_c3.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_8gpu_c3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['metric'],'| n',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'| bm25',round(d['kernels']['bm25_ms'],2),'dense',round(d['kernels']['dense_ms'],2),'other',round(d['kernels']['other_ms'],2))"
