#!/bin/bash
# Quick GPU check: full GPU test-suite, op microbench at the student batch, one bench run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== microbench"; S4_BENCH_B=24 timeout 300 python tools/bench_ops.py all 5 2>&1 | tee gpurun_out/bench_ops_b24.log
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
