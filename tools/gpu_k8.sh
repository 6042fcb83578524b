#!/bin/bash
# K8 fused (zb_vocab_topk + beam_cand_kernel): parity tests, then the decode leg A/B and a per-kernel timeline.
#   gpurun --timeout 600 -- 'bash tools/gpu_k8.sh r02ac'
tag=${1:-k8}
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "vocab_topk or candidates" > gpurun_out/${tag}_k8_tests.log 2>&1
rc=$?
echo "k8 tests rc=$rc"; grep -E "passed|failed|^FAILED|^E  |Mismatched|Greatest" gpurun_out/${tag}_k8_tests.log | head -30 | cut -c1-300
if [ $rc -eq 0 ]; then
  {
    ZB_BEAM_FUSED=0 timeout 60 python tools/decode_ab.py
    ZB_BEAM_FUSED=1 timeout 60 python tools/decode_ab.py
    ZB_BEAM_FUSED=0 timeout 60 python tools/decode_ab.py
    ZB_BEAM_FUSED=1 timeout 60 python tools/decode_ab.py
  } > gpurun_out/${tag}_decode_ab.jsonl 2> gpurun_out/${tag}_decode_ab.err
  python - <<PY
import json
for l in open("gpurun_out/${tag}_decode_ab.jsonl"):
    try:
        d = json.loads(l); print("%-40s %8.0f tok/s  %.4f ms/step" % (d["switches"], d["value"], d["ms_per_step"]))
    except Exception:
        print("bad line", l[:100])
PY
  tail -3 gpurun_out/${tag}_decode_ab.err
  ZB_BEAM_FUSED=1 timeout 120 python tools/trace_step.py --decode > gpurun_out/${tag}_decode_timeline_fused.log 2>&1
  grep -E "steps|gemm2|beam|idle" gpurun_out/${tag}_decode_timeline_fused.log | head -12 | cut -c1-200
fi
