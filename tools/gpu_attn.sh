#!/bin/bash
# attention kernels: parity, isolated timing (tcgen05 vs mma.sync), in-kernel timeline of CTA 0
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/${tag}_attn_tests.log 2>&1
echo "attention tests rc=$?"; grep -E "passed|failed|^FAILED|Mismatched|Greatest" gpurun_out/${tag}_attn_tests.log | head -40 | cut -c1-250
timeout 120 python tools/attn_bench.py --trace 2> gpurun_out/${tag}_attn_trace.log > /dev/null
timeout 200 python tools/attn_bench.py > gpurun_out/${tag}_attn_bench.jsonl 2>&1
cat gpurun_out/${tag}_attn_bench.jsonl | cut -c1-400
grep -A3 "fwd lq=64 lk=64" gpurun_out/${tag}_attn_trace.log | head -4
grep -A3 "bwd lq=64 lk=64" gpurun_out/${tag}_attn_trace.log | head -4
