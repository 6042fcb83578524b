#!/bin/bash
# In-kernel timeline (ZB_GEMM_TRACE=1) of the CTA-pair GEMM at the training step's shapes; last launch of each.
mkdir -p gpurun_out
for shape in ${SHAPES:-"4096,512,512,0,1,1,0" "4096,512,512,0,0,0,0" "4096,1536,512,0,1,1,0" "4096,2048,512,0,1,3,0" "4096,512,2048,0,1,1,0" "4096,2048,512,0,0,0,0" "4096,512,2048,0,0,0,0" "4096,1024,512,0,1,1,0"}; do
  sh=${shape//,/ }
  ZB_GEMM_TRACE=1 timeout 60 tools/gemm_selftest --one $sh 2>&1 | grep trace | tail -2
  timeout 60 tools/gemm_selftest --one $sh 2>&1 | grep "time m"
done | tee gpurun_out/${1:-gemm2_trace}.log
