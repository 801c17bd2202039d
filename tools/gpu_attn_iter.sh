#!/bin/bash
# Attention iteration on a B200 box: parity tests of the attention kernels, microbench, event traces.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export S4_BENCH_B=${S4_BENCH_B:-24}
echo "== attention parity"; timeout 600 python -m pytest tests/test_tc_gpu.py -q -m gpu -p no:cacheprovider -k "attn or attention" 2>&1 | tail -5 | tee gpurun_out/attn_tests.log
echo "== microbench"; timeout 300 python tools/bench_ops.py attn_fwd 5 2>&1 | tee gpurun_out/attn_bench.log
timeout 300 python tools/bench_ops.py attn_bwd 5 2>&1 | tee -a gpurun_out/attn_bench.log
for w in ${TRACES:-fwd bwd}; do
  timeout 300 python tools/attn_trace.py $w > gpurun_out/trace_$w.txt 2>&1; head -3 gpurun_out/trace_$w.txt
done
