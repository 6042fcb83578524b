#!/bin/bash
tag=${1:-r02u}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -q -k "split_k or add_ln or decode or beam" > gpurun_out/${tag}_decode_tests.log 2>&1
echo "decode tests rc=$?"; grep -E "passed|failed|^FAILED|Mismatched|Greatest" gpurun_out/${tag}_decode_tests.log | head -20 | cut -c1-250
{
  ZB_DECODE_SPLITK=0 timeout 60 python tools/decode_ab.py
  timeout 60 python tools/decode_ab.py
  ZB_DECODE_SPLITK=0 timeout 60 python tools/decode_ab.py
  timeout 60 python tools/decode_ab.py
} > gpurun_out/${tag}_decode_ab.jsonl 2> gpurun_out/${tag}_decode_ab.err
python - <<PY
import json
for l in open("gpurun_out/${tag}_decode_ab.jsonl"):
    try:
        d = json.loads(l); print("%-60s %8.0f tok/s  %.3f ms/step" % (d["switches"], d["value"], d["ms_per_step"]))
    except Exception:
        print("bad line", l[:100])
PY
tail -3 gpurun_out/${tag}_decode_ab.err
