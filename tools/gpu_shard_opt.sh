#!/bin/bash
# Validation + A/B of the fused gradient step (zb_shard_adam, ZB_SHARD_OPT) on an N-GPU box:
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_shard_opt.sh 2 r02'
# (1) kernel parity on one device with emulated ranks + the two-process training parity test, (2) bench.py at N ranks:
# NCCL all-reduce + replicated Adam (default) vs the fused step over the NVSwitch multicast mapping vs unicast peers.
n=${1:-2}; tag=${2:-shard}
mkdir -p gpurun_out
ZB_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests/test_shard_opt_gpu.py -m gpu -q -x > gpurun_out/${tag}_shard_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_shard_tests.log
tail -5 gpurun_out/${tag}_shard_tests.log
if grep -q "^rc=0" gpurun_out/${tag}_shard_tests.log; then
  for mode in 0 1 p2p; do
    ZB_SHARD_OPT=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --no-decode --no-extra \
      > gpurun_out/${tag}_bench_n${n}_shard_${mode}.json 2> gpurun_out/${tag}_bench_n${n}_shard_${mode}.err
    cut -c1-160 gpurun_out/${tag}_bench_n${n}_shard_${mode}.json
  done
fi
