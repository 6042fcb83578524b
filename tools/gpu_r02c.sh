#!/bin/bash
# round 2, visit c: the general tcgen05 attention kernels (attention_tc.cu) — parity, then the step timeline / bench
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > gpurun_out/${tag}_attn_tests.log 2>&1
echo "attention tests rc=$?"; tail -25 gpurun_out/${tag}_attn_tests.log | cut -c1-300
if grep -q "failed\|error" gpurun_out/${tag}_attn_tests.log; then
  # every case on its own so that one failure does not hide the others
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention_tcgen05" > gpurun_out/${tag}_attn_tests_all.log 2>&1
  grep -E "PASS|FAIL|passed|failed|Mismatch|Greatest" gpurun_out/${tag}_attn_tests_all.log | head -60 | cut -c1-250
  exit 1
fi
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 120 python tools/trace_step.py --steps 3 --csv gpurun_out/${tag}_timeline.csv > gpurun_out/${tag}_timeline.log 2>&1
grep -E "span|fa" gpurun_out/${tag}_timeline.log
ZB_ATTN_TC=0 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_mma.json 2>/dev/null
timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_tc.json 2>/dev/null
cut -c1-200 gpurun_out/${tag}_bench_mma.json gpurun_out/${tag}_bench_tc.json
