#!/bin/bash
# round-2 ncu evidence: (1) launch list of the bench command (gpu__time_duration), (2) --set full captures of the new
# kernels: tcgen05 attention fwd / bwd, the K6 GEMM passes, and the dominant GEMM next to them.  Raw pages are exported
# to CSV on the box (gpurun_out/ is capped at 64 MiB); only the attention report travels back as .ncu-rep.
tag=${1:-r02w}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-decode --no-extra > gpurun_out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${tag}_launches.csv)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fwd_tc_kernel|bwd_tc_kernel" --launch-skip 40 -c 4 -f \
  -o gpurun_out/${tag}_attention_tc_full python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-decode --no-extra \
  > gpurun_out/${tag}_ncu_attn.log 2>&1
echo "attention capture rc=$?"
ncu -i gpurun_out/${tag}_attention_tc_full.ncu-rep --page raw --csv > gpurun_out/${tag}_attention_tc_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k regex:"gemm2_bf16_tcgen05" --launch-skip 300 -c 40 -f \
  -o /tmp/${tag}_gemm2_full python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-decode --no-extra \
  > gpurun_out/${tag}_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu -i /tmp/${tag}_gemm2_full.ncu-rep --page raw --csv > gpurun_out/${tag}_gemm2_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
