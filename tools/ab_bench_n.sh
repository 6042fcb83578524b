#!/bin/bash
# tools/ab_bench_n.sh N "VAR=1" ... : like ab_bench.sh under torchrun with N ranks
n=$1; shift
for rep in 1 2; do
  for v in "" "$@"; do
    out=$(env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 40 --warmup 5 2>/dev/null | grep '^{' | tail -1)
    echo "[$rep] N=$n ${v:-default}: $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.0f tok/s  %.3f ms/step" % (d["value"], d["ms_per_step"]))')"
  done
done | tee gpurun_out/ab_bench_n$n.log
