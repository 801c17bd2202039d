#!/bin/bash
# ncu --set full captures of the attention kernels at the student batch (B=24) + op microbench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export S4_BENCH_B=${S4_BENCH_B:-24}
echo "== op microbench B=$S4_BENCH_B"; timeout 600 python tools/bench_ops.py all 5 2>&1 | tee gpurun_out/bench_ops_b$S4_BENCH_B.log
echo "== ncu attn fwd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 2 -f \
  -o gpurun_out/prof_attnf python tools/bench_ops.py attn_fwd 1 > gpurun_out/ncu_attnf.log 2>&1
tail -2 gpurun_out/ncu_attnf.log
echo "== ncu attn bwd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 1 -c 1 -f \
  -o gpurun_out/prof_attnb python tools/bench_ops.py attn_bwd 1 > gpurun_out/ncu_attnb.log 2>&1
tail -2 gpurun_out/ncu_attnb.log
ls -la gpurun_out/*.ncu-rep
