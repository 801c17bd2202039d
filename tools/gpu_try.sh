#!/bin/bash
# Quick GPU iteration: selected tests (own processes, short timeouts) + op micro-benchmarks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/tests
for t in "$@"; do
  case "$t" in
    bench:*) echo "== bench_ops ${t#bench:}"; timeout 300 python tools/bench_ops.py ${t#bench:} 10 2>&1 | tail -60 | tee -a gpurun_out/try_bench.log ;;
    *) echo "== pytest $t"; timeout 300 python -m pytest "$t" -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15 ;;
  esac
done
