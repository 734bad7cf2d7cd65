#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "dense_mma" > gpurun_out/pytest_mma.log 2>&1; echo "pytest mma exit $?"; tail -3 gpurun_out/pytest_mma.log
timeout 900 python scripts/bench_kernels.py 10000000 > gpurun_out/kernels_10m.json 2> gpurun_out/kernels_10m.err; echo "kernels exit $?"; cat gpurun_out/kernels_10m.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)"
tail -3 gpurun_out/kernels_10m.err
