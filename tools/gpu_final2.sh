#!/bin/bash
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? stdout lines=$(wc -l < gpurun_out/${tag}_bench.json)"; python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value %.0f tok/s  %.3f ms/step  e2e %.0f  roofline %.3f traffic %.3g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"]))
print("decode", d["decode"]["value"], d["decode"]["ms_per_step"], "clocks", d["clocks"])
PY
