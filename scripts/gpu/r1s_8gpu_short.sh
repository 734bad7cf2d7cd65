#!/bin/bash
# 8-GPU box, final kernels: headline c3 at 8 / 4 / 2 GPUs + c5
mkdir -p gpurun_out
run() { n=$1; name=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $? =="; tail -n 1 gpurun_out/$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['metric'],'| n',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'| bm25',round(d['kernels']['bm25_ms'],2),'dense',round(d['kernels']['dense_ms'],2),'other',round(d['kernels']['other_ms'],2))" || tail -n 8 gpurun_out/$name.log; }
run 8 bench_8gpu_c3 --steps 20 --warmup 3
run 4 bench_4gpu_c3 --steps 10 --warmup 3
run 2 bench_2gpu_c3 --steps 10 --warmup 3
run 8 bench_8gpu_c5 --workload c5 --steps 5 --warmup 3
