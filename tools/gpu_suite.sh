#!/bin/bash
# full GPU suite + step timeline + bench with the tcgen05 attention (default) and with ZB_ATTN_TC=0
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/${tag}_tests.log | cut -c1-300
ZB_ATTN_TC=0 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_mma.json 2>/dev/null
timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_tc.json 2>/dev/null
ZB_ATTN_TC=0 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_mma2.json 2>/dev/null
timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_tc2.json 2>/dev/null
cut -c1-200 gpurun_out/${tag}_bench_mma.json gpurun_out/${tag}_bench_tc.json gpurun_out/${tag}_bench_mma2.json gpurun_out/${tag}_bench_tc2.json
