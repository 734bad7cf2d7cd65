#!/bin/bash
for st in 7 6 5 4 3; do echo "stages $st"; RAGB_MMA_STAGES=$st MMA_VARIANTS=3 MMA_KS=50 timeout 600 python scripts/bench_mma.py 10000000 2>&1 | tail -1; done
