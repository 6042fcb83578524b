#!/bin/bash
# Runs the torch-free GEMM self-test per operand layout (a separate process each, so one trap does not hide the rest).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for a in 0 1; do for b in 0 1; do
  echo "=== layout A_MN=$a B_MN=$b"; timeout 120 tools/gemm_selftest --layout $a $b 2>&1 | tail -30
done; done | tee gpurun_out/gemm_selftest.log
timeout 300 tools/gemm_selftest --time 2>&1 | grep -E "time|SELFTEST" | tee gpurun_out/gemm_time.log
