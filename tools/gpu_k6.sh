#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "vocab_ce or softmax_ce or gemm" > gpurun_out/${tag}_k6_tests.log 2>&1
echo "k6 kernel tests rc=$?"; grep -E "passed|failed|^FAILED|Mismatched|Greatest|Error|assert" gpurun_out/${tag}_k6_tests.log | head -30 | cut -c1-250
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/${tag}_tests.log | cut -c1-300
ZB_FUSED_CE=0 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_unfused.json 2>/dev/null
timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_fused.json 2>/dev/null
cut -c1-200 gpurun_out/${tag}_bench_unfused.json gpurun_out/${tag}_bench_fused.json
timeout 120 python tools/trace_step.py --steps 3 --csv gpurun_out/${tag}_timeline.csv > gpurun_out/${tag}_timeline.log 2>&1
grep -E "span|n/step" gpurun_out/${tag}_timeline.log | head -24
