#!/bin/bash
# One GPU-box visit: parity tests, bench line, in-situ timeline, ncu launch list, one full ncu capture of the GEMM.
# usage: tools/gpu_round.sh TAG   (outputs under gpurun_out/TAG_*)
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 120 python tools/trace_step.py --steps 3 --csv gpurun_out/${tag}_timeline.csv > gpurun_out/${tag}_trace.log 2>&1
if [ -z "$SKIP_NCU" ]; then
[ -z "$SKIP_LIST" ] && timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-decode --no-extra > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-gemm2_bf16_tcgen05} --launch-skip ${NCU_SKIP:-60} -c ${NCU_C:-6} -f -o gpurun_out/${tag}_full \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-decode --no-extra > gpurun_out/${tag}_ncu_full.log 2>&1
fi
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench.json | cut -c1-600; tail -25 gpurun_out/${tag}_trace.log
