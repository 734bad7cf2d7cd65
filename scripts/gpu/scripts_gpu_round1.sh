#!/bin/bash
# First GPU pass: parity tests in isolated processes (a faulting kernel poisons its own process only),
# then smoke, then a reduced-size bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() {
  name=$1; shift
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider "$@" > gpurun_out/pytest_$name.log 2>&1
  echo "== $name exit $? ==" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/pytest_$name.log
}
run select -k "topk_rows or topk_merge or hybrid_fuse or error_behaviour"
run bm25 -k "bm25"
run gemv -k "dense_gemv"
run mma0 -k "dense_mma and 0-"
run mma1 -k "dense_mma and 1-"
run router -k "router"
run mc -k "mc_dropout"
run e2e -k "end_to_end or row_sharded or dropin or gemv_equals"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --passages 1000000 --steps 5 --warmup 3 > gpurun_out/bench_1m.log 2>&1; echo "== bench1m exit $? ==" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/bench_1m.log
cat gpurun_out/summary.txt
