#!/bin/bash
# One GPU round on a B200 box: the driver's test command, smoke, bench, and the ncu launch list.
# Logs land in gpurun_out/.  Usage: tools/gpu_round.sh [tests|bench|ncu|all]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == all || $what == tests ]]; then
  echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
  echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  echo "== bench"; timeout 1200 python bench.py --steps ${STEPS:-5} --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
  tail -5 gpurun_out/bench.err
fi
if [[ $what == all || $what == ncu ]]; then
  echo "== ncu launch list"
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prof \
    > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log
  wc -l gpurun_out/launches.csv
fi
