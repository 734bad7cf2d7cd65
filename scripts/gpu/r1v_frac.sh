#!/bin/bash
for f in 8 16 32 64; do echo "dense table for df >= N/$f"; RAGB_DENSE_MIN_FRACTION=$f timeout 600 python scripts/bench_bm25.py 10000000 50 2>&1 | tail -1; done
