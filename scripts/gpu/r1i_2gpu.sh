#!/bin/bash
# 2-GPU check of the sharded path: headline workload, c4 (MC-Dropout), full-fusion
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $? =="; tail -n 1 gpurun_out/$name.log | cut -c1-900; }
run bench_2gpu_c3 --steps 5 --warmup 3
run bench_2gpu_c4 --steps 5 --warmup 3 --workload c4
run bench_2gpu_ff --steps 3 --warmup 3 --mode full-fusion
