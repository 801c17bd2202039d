#!/bin/bash
# ncu --set full captures of the three tensor-core kernels at the student batch (B=24).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export S4_BENCH_B=24
for spec in "gemm:fc1:gemm_tc_kernel:1" "attnf:attn_fwd:attn_fwd_kernel:1" "attnb:attn_bwd:attn_bwd_kernel:1"; do
  IFS=: read tag what pat skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f \
    -o gpurun_out/prof_$tag python tools/bench_ops.py $what 2 > gpurun_out/ncu_$tag.log 2>&1
  tail -1 gpurun_out/ncu_$tag.log | cut -c1-100
done
ls -la gpurun_out/*.ncu-rep
