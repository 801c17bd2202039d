#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / TMEM / TMA kernels:
# memcheck on the GEMM / conv / attention / head kernel tests, racecheck + synccheck on the attention and
# GEMM kernels (shared-memory hazards between the producer / MMA / epilogue roles).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=compute-sanitizer
run() {  # name tool filter
  timeout ${4:-600} $S --tool $2 --print-limit 5 --error-exitcode 7 \
    python -m pytest tests/test_tc_gpu.py -x -q -m gpu -p no:cacheprovider -k "$3" > gpurun_out/sanitize_$1.log 2>&1
  echo "$1: rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$1.log | tr '\n' ' ')"
}
run memcheck_gemm memcheck "tc_gemm_kmajor_pairs or layouts_and_splitk" 500
run memcheck_attn memcheck "tc_attention" 500
run racecheck_attn racecheck "tc_attention_fwd or tc_attention_bwd" 700
run synccheck_attn synccheck "tc_attention_fwd or tc_attention_bwd" 500
run racecheck_gemm racecheck "tc_gemm_kmajor_pairs" 500
