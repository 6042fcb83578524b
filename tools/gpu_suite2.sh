#!/bin/bash
# attention kernel tests first (fast failure), then the whole suite, then the full bench line with all legs
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/${tag}_attn_tests.log 2>&1
echo "attention rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/${tag}_attn_tests.log | head -20 | cut -c1-250
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 600 python bench.py --steps 50 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value %.0f tok/s  %.3f ms/step  e2e %.0f  roofline %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
    print("cpu_baseline", d["cpu_baseline"])
    print("decode", {k: d["decode"][k] for k in ("value", "ms_per_step")} if d.get("decode") else None)
    for k, v in (d.get("legs") or {}).items():
        print(k, json.dumps(v)[:600])
except Exception as e:
    print("bad bench line", e); print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
