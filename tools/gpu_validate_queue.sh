#!/bin/bash
# First GPU-box visit after a stretch of CPU-only work: (1) the default parity suite, (2) the opt-in parity tests of
# every kernel / switch written without a GPU, (3) the decode A/B over those switches,
# (4) ncu --set full of the two kernels whose time is not explained yet.  Outputs under gpurun_out/TAG_*.
tag=${1:-queue}
mkdir -p gpurun_out
t0=$SECONDS
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$? in $((SECONDS-t0)) s" >> gpurun_out/${tag}_tests.log
# one process per group: a kernel that traps (every mbarrier wait is bounded) poisons its CUDA context, not the others
: > gpurun_out/${tag}_tests_unvalidated.log
for grp in beam_part gemm_bm64 fused_small opt_in_decode attention_tcgen05 batched_memory "gumbel or noise_beam" "shard_adam_kernel or shard_adam_rejects" vocabulary_size bucketed_graph; do
  t0=$SECONDS
  echo "=== $grp" >> gpurun_out/${tag}_tests_unvalidated.log
  ZB_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests -m gpu -q -k "$grp" >> gpurun_out/${tag}_tests_unvalidated.log 2>&1
  echo "=== $grp rc=$? in $((SECONDS-t0)) s" >> gpurun_out/${tag}_tests_unvalidated.log
done
{
  timeout 60 python tools/decode_ab.py
  ZB_BEAM_PARTS=1 timeout 60 python tools/decode_ab.py
  ZB_GEMM_BM64=1 timeout 60 python tools/decode_ab.py
  ZB_DECODE_FUSED_SMALL=1 timeout 60 python tools/decode_ab.py
  ZB_BEAM_PARTS=1 ZB_GEMM_BM64=1 ZB_DECODE_FUSED_SMALL=1 timeout 60 python tools/decode_ab.py
} > gpurun_out/${tag}_decode_ab.jsonl 2> gpurun_out/${tag}_decode_ab.err
# training bench: the default first (the baseline of every A/B below), then each switch whose parity test passed
timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_default.json 2>/dev/null
if grep -q "=== attention_tcgen05 rc=0" gpurun_out/${tag}_tests_unvalidated.log; then
  ZB_ATTN_TC=1 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_attn_tc.json 2>/dev/null
  cut -c1-200 gpurun_out/${tag}_bench_default.json gpurun_out/${tag}_bench_attn_tc.json
fi
# batched memory projection: training bench under the switch (its parity test ran above)
if grep -q "=== batched_memory rc=0" gpurun_out/${tag}_tests_unvalidated.log; then
  ZB_BATCH_MEM_PROJ=1 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_batchmem.json 2>/dev/null
  cut -c1-200 gpurun_out/${tag}_bench_batchmem.json
fi
# L2 -> SM ingest cap and TMA multicast at cluster sizes 2 / 4 / 8 (decides whether shared-A clusters are worth building)
timeout 60 tools/tma_l2_probe > gpurun_out/${tag}_tma_l2_probe.log 2>&1; tail -12 gpurun_out/${tag}_tma_l2_probe.log
# GEMM tile width: byte-weighted wave rule / 256-wide tiles everywhere (validated kernels, only the choice differs)
ZB_GEMM2_TILE_MODEL=l2 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_tile_l2.json 2>/dev/null
ZB_GEMM2_BN=256 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_bn256.json 2>/dev/null
cut -c1-200 gpurun_out/${tag}_bench_tile_l2.json gpurun_out/${tag}_bench_bn256.json
# add+LN backward with 16 rows per CTA: parity test under the switch, then the training bench
ZB_LN1P_WARPS=16 timeout 60 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "add_ln" > gpurun_out/${tag}_tests_ln16.log 2>&1
if tail -1 gpurun_out/${tag}_tests_ln16.log | grep -q passed && ! grep -q failed gpurun_out/${tag}_tests_ln16.log; then
  ZB_LN1P_WARPS=16 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_ln16.json 2>/dev/null
  cut -c1-200 gpurun_out/${tag}_bench_ln16.json
fi
# add+LN backward with 4 rows per warp, 8 warps per CTA (column partials in registers): same protocol
ZB_LN1P_WARPS=8 timeout 60 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "add_ln" > gpurun_out/${tag}_tests_ln8.log 2>&1
if tail -1 gpurun_out/${tag}_tests_ln8.log | grep -q passed && ! grep -q failed gpurun_out/${tag}_tests_ln8.log; then
  ZB_LN1P_WARPS=8 timeout 100 python bench.py --no-cpu-baseline --no-decode --no-extra --steps 50 > gpurun_out/${tag}_bench_ln8.json 2>/dev/null
  cut -c1-200 gpurun_out/${tag}_bench_ln8.json
fi
if [ -z "$SKIP_NCU" ]; then
ZB_DECODE_GRAPH=0 timeout 90 ncu --set full --clock-control none --import-source on -k regex:"beam_row|beam_part" \
  --launch-skip 70 -c 2 -f -o gpurun_out/${tag}_beam_full python tools/decode_ab.py 1 > gpurun_out/${tag}_ncu_beam.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"add_ln_bwd_1pass" --launch-skip 40 -c 2 -f \
  -o gpurun_out/${tag}_lnbwd_full python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-decode --no-extra \
  > gpurun_out/${tag}_ncu_lnbwd.log 2>&1
fi
grep -E "passed|failed|error" gpurun_out/${tag}_tests.log | tail -2
grep -E "^=== .* rc=|passed|failed" gpurun_out/${tag}_tests_unvalidated.log; grep -E "^FAILED|^ERROR" gpurun_out/${tag}_tests_unvalidated.log | head -30
python - <<PY
import json
for l in open("gpurun_out/${tag}_decode_ab.jsonl"):
    try:
        d = json.loads(l); print("%-100s %8.0f tok/s  %.3f ms/step" % (d["switches"], d["value"], d["ms_per_step"]))
    except Exception:
        print("bad line", l[:100])
PY
tail -3 gpurun_out/${tag}_decode_ab.err
python tools/summarize_queue.py ${tag}
