#!/bin/bash
mkdir -p gpurun_out
timeout 60 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit $?"; tail -n 2 gpurun_out/sanitize_plain.log
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"; tail -n 6 gpurun_out/memcheck.log
