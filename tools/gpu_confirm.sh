#!/bin/bash
# Short GPU-box visit: parity tests, bench line, GEMM DRAM-traffic pass (ncu, 3 metrics), decode timeline.
# usage: tools/gpu_confirm.sh TAG   (outputs under gpurun_out/TAG_*)
tag=${1:-confirm}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
t0=$SECONDS
timeout ${T_TESTS:-300} python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$? in $((SECONDS-t0)) s" >> gpurun_out/${tag}_tests.log
t0=$SECONDS
timeout 200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? in $((SECONDS-t0)) s" >> gpurun_out/${tag}_bench.err
timeout 60 python tools/trace_step.py --decode --csv gpurun_out/${tag}_decode_timeline.csv > gpurun_out/${tag}_decode_trace.log 2>&1
timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:gemm2_bf16_tcgen05 --csv --log-file gpurun_out/${tag}_gemm_traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-decode --no-extra > gpurun_out/${tag}_ncu_traffic.log 2>&1
tail -4 gpurun_out/${tag}_tests.log; cut -c1-400 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err; head -30 gpurun_out/${tag}_decode_trace.log
