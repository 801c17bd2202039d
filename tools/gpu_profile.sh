#!/bin/bash
# ncu passes on a B200 box: (1) launch list of one bench step, (2) --set full captures of the
# top kernels at the train step's shapes (via tools/bench_ops.py).  Output: gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== op microbench"; timeout 600 python tools/bench_ops.py all 10 2>&1 | tee gpurun_out/bench_ops.log
echo "== ncu launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prof \
  > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/launches.csv
for spec in "gemm:fc1:gemm_tc_kernel" "attnf:attn_fwd:attn_fwd_kernel" "attnb:attn_bwd:attn_bwd_kernel"; do
  IFS=: read tag what pat <<< "$spec"
  echo "== ncu --set full $tag"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 2 -f \
    -o gpurun_out/prof_$tag python tools/bench_ops.py $what 3 > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
done
ls -la gpurun_out/
