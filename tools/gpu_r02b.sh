#!/bin/bash
# round 2, visit b: new defaults (batched memory projection, byte-weighted tile rule, cluster beam kernel, fused small
# decode kernels), the parity tests at the BASELINE shapes with their achieved errors, step timeline with the
# mma.sync and the tcgen05 64-token attention kernels
tag=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_smi.txt
timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py -m gpu -q -s > gpurun_out/${tag}_baseline_shapes.log 2>&1
echo "baseline shapes rc=$?"; grep -E "^C[2345]:|identical|passed|failed" gpurun_out/${tag}_baseline_shapes.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_baseline_shapes_gpu.py > gpurun_out/${tag}_tests.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/${tag}_tests.log
timeout 120 python tools/trace_step.py --steps 3 --csv gpurun_out/${tag}_timeline_default.csv > gpurun_out/${tag}_timeline_default.log 2>&1
ZB_ATTN_TC=1 timeout 120 python tools/trace_step.py --steps 3 --csv gpurun_out/${tag}_timeline_attn_tc.csv > gpurun_out/${tag}_timeline_attn_tc.log 2>&1
grep -E "span|fa::|add_ln|colsum" gpurun_out/${tag}_timeline_default.log gpurun_out/${tag}_timeline_attn_tc.log
timeout 200 python bench.py --steps 50 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-400 gpurun_out/${tag}_bench.json
