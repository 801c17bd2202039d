#!/bin/bash
# ncu SpeedOfLight + memory sections for the HBM-bound kernels of one bench step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PAT='regex:ce_ncr_kernel|colsum_vec|ln_fwd_kernel|ln_bwd_kernel|bn_relu_upsample|cls_bwd_apply|cls_bwd_reduce|cls_fwd_kernel|cls_upsample|bn_bwd_apply|pseudo_label|sgd_multi|ema_multi|patchify|pack_conv|attn_bwd_delta|attn_bwd_dq'
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__warps_active.avg.per_cycle_active \
  --clock-control none -k "$PAT" -s ${SKIP:-900} -c ${COUNT:-330} --csv --page raw \
  --log-file gpurun_out/elem_raw.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prof > gpurun_out/ncu_elem.log 2>&1
tail -2 gpurun_out/ncu_elem.log | cut -c1-200
wc -l gpurun_out/elem_raw.csv
