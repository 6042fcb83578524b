"""Beam-4 decode leg of bench.py alone (BASELINE configs[2]), for A/B runs of the decode-path switches:
ZB_DECODE_ATTN (lq = 1 attention kernel), ZB_BEAM_ROWS (row-parallel beam step), ZB_DECODE_SPEC (host one step
ahead), ZB_SKINNY_GEMM (small-tile GEMM for <= 384 rows).  Prints one JSON line tagged with the switches set."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    out = bench.decode_bench("cuda:0", batches=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
    out["switches"] = {k: v for k, v in os.environ.items() if k.startswith("ZB_")}
    print(json.dumps(out))
