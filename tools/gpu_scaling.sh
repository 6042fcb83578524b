#!/bin/bash
# Data-parallel A/B on an N-GPU box:   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_scaling.sh N tag'
#   ZB_ENC_BUCKETS=1 (round-1 two-bucket scheme) vs 3 / 6 encoder groups + late embedding bucket under Adam,
#   ZB_SHARD_OPT=1 / p2p (fused reduce-scatter + Adam + all-gather kernel), single-GPU reference point.
n=${1:-2}; tag=${2:-scale}
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
    print("N=${n} %-16s %9.0f tok/s  %.3f ms/step" % ("${label}", d["value"], d["ms_per_step"]))
except Exception as e:
    print("N=${n} ${label}: failed", e); print(open("gpurun_out/${tag}_n${n}_${label}.err").read()[-1500:])
PY
}
timeout 100 python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-decode --no-extra > gpurun_out/${tag}_n1.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_n1.json').read()); print('N=1 %9.0f tok/s  %.3f ms/step' % (d['value'], d['ms_per_step']))"
if [ "$n" = "2" ]; then
  timeout 600 python -m pytest tests/test_shard_opt_gpu.py -m gpu -q -x > gpurun_out/${tag}_shard_tests.log 2>&1
  echo "shard tests rc=$?"; tail -2 gpurun_out/${tag}_shard_tests.log | cut -c1-200
fi
run buckets1 ZB_ENC_BUCKETS=1
run buckets3 ZB_ENC_BUCKETS=3
run buckets6 ZB_ENC_BUCKETS=6
run shard_mc ZB_SHARD_OPT=1
run shard_p2p ZB_SHARD_OPT=p2p
run buckets3_b ZB_ENC_BUCKETS=3
