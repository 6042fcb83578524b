#!/bin/bash
# A/B the training step under environment variants:  tools/ab_bench.sh "ZB_X=1" "ZB_Y=1 ZB_Z=0" ...
# (the empty variant "" is the default build); prints value / ms_per_step of each, twice, interleaved.
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "" "$@"; do
    out=$(env $v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-decode --no-extra 2>/dev/null | tail -1)
    echo "[$rep] ${v:-default}: $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.0f tok/s  %.3f ms/step  gemm %.3f ms" % (d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"]))')"
  done
done | tee gpurun_out/ab_bench.log
