#!/usr/bin/env bash
# One parameterised runner for everything that needs the B200 box (replaces the per-experiment shell scripts of
# round 1).  Run it through gpurun from the repository root, e.g.
#
#   gpurun --timeout 1500 -- 'bash scripts/gpu_suite.sh tag=r02a tests smoke bench "bench:--workload c2" launches ncu:bm25_kernel'
#
# Steps (executed in order; every step writes under gpurun_out/<tag>_*; a failing step does not stop the rest):
#   tag=<name>                 prefix of the output files (default: run)
#   env:<VAR=value>            export a variable for the following steps (tuning knobs)
#   tests[:<pytest -k expr>]   python -m pytest tests -m gpu -x -q [-k expr]
#   smoke                      python -c "import __graft_entry__ as g; g.smoke()"
#   bench[:<extra args>]       python bench.py <extra args>       (JSON line -> <tag>_bench<i>.json)
#   torchrun:<N>[:<args>]      the driver's multi-GPU launch of bench.py on N GPUs
#   launches[:<extra args>]    ncu launch list (gpu__time_duration.sum) of bench.py --steps 2 --warmup 1 <args>,
#                              restricted to the NVTX range bench_timed
#   ncu:<kernel regex>[:<args>] one `ncu --set full` capture of the first matching launch(es) inside bench_timed (env
#                              NCU_COUNT launches, default 1: the dense kernel runs twice per step, sampled prefix then the
#                              seeded search), exported as <tag>_ncu_<regex>.ncu-rep plus a raw CSV page
#   py:<script and args>       python <script and args>           (scripts/*.py micro-benchmarks)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=run
i=0
for step in "$@"; do
  i=$((i + 1))
  name=${step%%:*}
  rest=""
  [[ "$step" == *:* ]] && rest=${step#*:}
  case "$name" in
    tag=*) tag=${name#tag=} ;;
    env) export "$rest"; echo "[suite] export $rest" ;;
    tests)
      if [[ -n "$rest" ]]; then python -m pytest tests -m gpu -x -q -k "$rest" > "gpurun_out/${tag}_tests${i}.log" 2>&1
      else python -m pytest tests -m gpu -x -q > "gpurun_out/${tag}_tests${i}.log" 2>&1; fi
      echo "[suite] tests rc=$? $(tail -n 1 "gpurun_out/${tag}_tests${i}.log")" ;;
    smoke)
      python -c "import __graft_entry__ as g; g.smoke()" > "gpurun_out/${tag}_smoke.log" 2>&1
      echo "[suite] smoke rc=$? $(tail -n 1 "gpurun_out/${tag}_smoke.log")" ;;
    bench)
      # shellcheck disable=SC2086
      python bench.py $rest > "gpurun_out/${tag}_bench${i}.json" 2> "gpurun_out/${tag}_bench${i}.err"
      echo "[suite] bench $rest rc=$? $(head -c 600 "gpurun_out/${tag}_bench${i}.json")" ;;
    torchrun)
      n=${rest%%:*}; args=""; [[ "$rest" == *:* ]] && args=${rest#*:}
      # shellcheck disable=SC2086
      python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus "$n" $args > "gpurun_out/${tag}_bench${i}_${n}gpu.json" 2> "gpurun_out/${tag}_bench${i}_${n}gpu.err"
      echo "[suite] torchrun $n $args rc=$? $(grep -m1 '^{' "gpurun_out/${tag}_bench${i}_${n}gpu.json" | head -c 600)" ;;
    launches)
      # shellcheck disable=SC2086
      ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bench_timed/" -c 400 --csv \
        --log-file "gpurun_out/${tag}_launches${i}.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 $rest \
        > "gpurun_out/${tag}_launches${i}.log" 2>&1
      echo "[suite] launches rc=$?" ;;
    ncu)
      kern=${rest%%:*}; args=""; [[ "$rest" == *:* ]] && args=${rest#*:}
      # shellcheck disable=SC2086
      ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "bench_timed/" -k "regex:${kern}" -c "${NCU_COUNT:-1}" \
        -o "gpurun_out/${tag}_ncu_${kern}" -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 $args \
        > "gpurun_out/${tag}_ncu_${kern}.log" 2>&1
      rc=$?
      ncu -i "gpurun_out/${tag}_ncu_${kern}.ncu-rep" --page raw --csv > "gpurun_out/${tag}_ncu_${kern}_raw.csv" 2>/dev/null
      echo "[suite] ncu $kern rc=$rc" ;;
    py)
      # shellcheck disable=SC2086
      python $rest > "gpurun_out/${tag}_py${i}.log" 2>&1
      echo "[suite] py $rest rc=$? $(tail -n 3 "gpurun_out/${tag}_py${i}.log")" ;;
    *) echo "[suite] unknown step: $step" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > "gpurun_out/${tag}_smi.txt" 2>&1
echo "[suite] done"
