#!/bin/bash
# Runs the GPU test-suite on a B200 box; every group in its own process (a trapped kernel
# poisons the CUDA context) with a timeout; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/tests
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/kernels.log
echo "== tc (one process per test)"
: > gpurun_out/tc.log
for t in $(python -m pytest tests/test_tc_gpu.py --collect-only -q -m gpu 2>/dev/null | grep '::'); do
  out=$(timeout 240 python -m pytest "$t" -q -m gpu -p no:cacheprovider 2>&1 | tail -25)
  if echo "$out" | grep -q " passed"; then echo "PASS $t" | tee -a gpurun_out/tc.log
  else echo "FAIL $t" | tee -a gpurun_out/tc.log; echo "$out" | grep -E "Error|error|assert|rel err|trap|illegal" | head -8 | tee -a gpurun_out/tc.log; fi
done
echo "== step"; timeout 1200 python -m pytest tests/test_step_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/step.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
