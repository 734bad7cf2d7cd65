#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $? =="; tail -n 1 gpurun_out/$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['metric'],'| n',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'| bm25',round(d['kernels']['bm25_ms'],2),'dense',round(d['kernels']['dense_ms'],2),'other',round(d['kernels']['other_ms'],3))" || tail -n 8 gpurun_out/$name.log; }
run bench_2gpu_c3 --steps 10 --warmup 3
