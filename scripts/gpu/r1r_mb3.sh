#!/bin/bash
# A/B: bm25_kernel compiled for 4 (64 registers, small spills) vs 3 (80 registers, no spills) resident blocks per SM
PKG=$(ls -d efficient-rag*_b200)
echo "min blocks 4 (default)"; timeout 600 python scripts/bench_bm25.py 10000000 50 2>&1 | tail -1
timeout 600 python scripts/bench_bm25.py 1250000 50 2>&1 | tail -1
cp $PKG/libragb200.so /tmp/lib_default.so; cp $PKG/build/libragb200_mb3.so $PKG/libragb200.so
echo "min blocks 3"; timeout 600 python scripts/bench_bm25.py 10000000 50 2>&1 | tail -1
timeout 600 python scripts/bench_bm25.py 1250000 50 2>&1 | tail -1
cp /tmp/lib_default.so $PKG/libragb200.so
