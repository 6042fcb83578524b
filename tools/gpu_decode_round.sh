#!/bin/bash
# GPU-box visit for the decode path: parity tests, decode A/B over the path switches, decode timeline, bench line.
tag=${1:-dec}
mkdir -p gpurun_out
t0=$SECONDS
timeout 200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${tag}_tests.log 2>&1; rc=$?
echo "tests rc=$rc in $((SECONDS-t0)) s" >> gpurun_out/${tag}_tests.log
if [ $rc -ne 0 ]; then
  t0=$SECONDS
  ZB_SKINNY_SPLIT=0 timeout 120 python -m pytest tests -m gpu -q \
    > gpurun_out/${tag}_tests_nosplit.log 2>&1; echo "tests(no split) rc=$? in $((SECONDS-t0)) s" >> gpurun_out/${tag}_tests_nosplit.log
fi
{
  timeout 60 python tools/decode_ab.py
  ZB_SKINNY_SPLIT=0 timeout 60 python tools/decode_ab.py
  ZB_SKINNY_GEMM=0 timeout 60 python tools/decode_ab.py
  if [ -n "$FULL_AB" ]; then
  ZB_DECODE_SPEC=0 timeout 60 python tools/decode_ab.py
  ZB_BEAM_ROWS=0 timeout 60 python tools/decode_ab.py
  ZB_DECODE_ATTN=0 timeout 60 python tools/decode_ab.py
  ZB_DECODE_ATTN=0 ZB_BEAM_ROWS=0 ZB_DECODE_SPEC=0 ZB_SKINNY_GEMM=0 timeout 60 python tools/decode_ab.py
  fi
} > gpurun_out/${tag}_decode_ab.jsonl 2> gpurun_out/${tag}_decode_ab.err
timeout 60 python tools/trace_step.py --decode --csv gpurun_out/${tag}_decode_timeline.csv > gpurun_out/${tag}_decode_trace.log 2>&1
if [ -z "$SKIP_BENCH" ]; then
timeout 150 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
fi
grep -E "passed|failed|error" gpurun_out/${tag}_tests.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/${tag}_tests.log | head -20
[ -f gpurun_out/${tag}_tests_nosplit.log ] && tail -2 gpurun_out/${tag}_tests_nosplit.log
python - <<PY
import json
for l in open("gpurun_out/${tag}_decode_ab.jsonl"):
    try:
        d = json.loads(l); print("%-90s %8.0f tok/s  %.3f ms/step" % (d["switches"], d["value"], d["ms_per_step"]))
    except Exception as e:
        print("bad line", l[:100])
PY
tail -3 gpurun_out/${tag}_decode_ab.err; head -22 gpurun_out/${tag}_decode_trace.log | tail -20
