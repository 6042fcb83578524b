#!/bin/bash
n=${1:-2}; tag=${2:-scale2}
mkdir -p gpurun_out
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n --steps 60 --warmup 8 --no-cpu-baseline --no-decode --no-extra \
    > gpurun_out/${tag}_n${n}_${label}.json 2> gpurun_out/${tag}_n${n}_${label}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_n${n}_${label}.json") if l.startswith("{")][-1])
    print("N=${n} %-22s %9.0f tok/s  %.3f ms/step" % ("${label}", d["value"], d["ms_per_step"]))
except Exception as e:
    print("N=${n} ${label}: failed", e); print(open("gpurun_out/${tag}_n${n}_${label}.err").read()[-1500:])
PY
}
shift 2
while [ $# -gt 0 ]; do
  label=$1; shift
  envs=()
  while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
  [ "$1" = "--" ] && shift
  run "$label" "${envs[@]}"
done
