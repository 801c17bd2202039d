#!/bin/bash
# DRAM traffic of the tcgen05 GEMM / implicit-GEMM conv launches of one train step (for roofline.traffic).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:gemm_tc_kernel -s ${SKIP:-780} -c ${COUNT:-260} --csv --page raw \
  --log-file gpurun_out/gemm_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prof > gpurun_out/ncu_gemm_traffic.log 2>&1
tail -1 gpurun_out/ncu_gemm_traffic.log | cut -c1-120
wc -l gpurun_out/gemm_traffic.csv
