#!/bin/bash
# Last GPU-box visit of the round: parity tests, bench line, small-GEMM floor probe, ncu --set full of the decode kernels.
tag=${1:-final}
mkdir -p gpurun_out
t0=$SECONDS
timeout 70 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$? in $((SECONDS-t0)) s" >> gpurun_out/${tag}_tests.log
t0=$SECONDS
timeout 110 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? in $((SECONDS-t0)) s" >> gpurun_out/${tag}_bench.err
timeout 60 python tools/skinny_probe.py > gpurun_out/${tag}_skinny_probe.jsonl 2> gpurun_out/${tag}_skinny_probe.err
ZB_DECODE_GRAPH=0 timeout 80 ncu --set full --clock-control none --import-source on \
  -k regex:"beam_row|attn_decode|gemm2_bf16" --launch-skip 2660 -c 14 -f -o gpurun_out/${tag}_decode_full \
  python tools/decode_ab.py 1 > gpurun_out/${tag}_ncu.log 2>&1
grep -E "passed|failed|error" gpurun_out/${tag}_tests.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/${tag}_tests.log | head
cut -c1-300 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_skinny_probe.jsonl; tail -3 gpurun_out/${tag}_skinny_probe.err; tail -3 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/ | grep ${tag}
