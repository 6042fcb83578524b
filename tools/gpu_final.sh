#!/bin/bash
# end-of-round validation: the whole GPU suite, the full bench line (all legs), memcheck over the new kernels
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value %.0f tok/s  %.3f ms/step  e2e %.0f  roofline %.3f  launches/step %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("kernels_per_step")))
    print("cpu_baseline", d["cpu_baseline"])
    print("decode", {k: d["decode"][k] for k in ("value", "ms_per_step")} if d.get("decode") else None)
    for k, v in (d.get("legs") or {}).items():
        print(k, json.dumps(v)[:700])
except Exception as e:
    print("bad bench line", e); print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x \
  -k "(attention_tcgen05 and (4-8-64-64-False or 2-2-300-300 or 2-8-128-128-True)) or (relative_positions and 2-2-40-36) or (vocab_ce and 5-9-1000) or (rela and 2-2-40-36)" \
  > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/${tag}_memcheck.log | head -8 | cut -c1-200
