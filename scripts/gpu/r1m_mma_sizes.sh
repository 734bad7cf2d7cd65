#!/bin/bash
for pub in 4 2 1; do for p in 4 3; do echo "publish $pub pace $p"; RAGB_MMA_PUBLISH=$pub RAGB_MMA_PACE=$p MMA_VARIANTS=3 MMA_KS=50 timeout 600 python scripts/bench_mma.py 10000000 2>&1 | tail -1; done; done
RAGB_MMA_PUBLISH=2 MMA_VARIANTS=3 MMA_KS=50 timeout 600 python scripts/bench_mma.py 1250000 2>&1 | tail -1
MMA_VARIANTS=3 MMA_KS=50 timeout 600 python scripts/bench_mma.py 1250000 2>&1 | tail -1
