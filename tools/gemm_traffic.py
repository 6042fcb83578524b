"""profiles/gemm_dram_traffic.json from an ncu pass over the GEMM launches of one training step:

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:gemm2_bf16_tcgen05 --csv --log-file gpurun_out/gemm_traffic.csv \
      python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-decode
  python tools/gemm_traffic.py gpurun_out/gemm_traffic.csv 147

The last N launches in the log are the roofline leg's eager step (bench.gemm_roofline records exactly one step).
dram bytes per launch = (read + write) summed over those launches / N, next to the algorithmic operand bytes."""
import collections
import csv
import json
import os
import sys


def main(path, n):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        name = row["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        else:
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        per.setdefault(row["ID"], {})[name] = v
    ids = list(per)[-n:]
    rd = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in ids)
    wr = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in ids)
    us = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in ids)
    out = {"launches": len(ids), "dram_bytes_per_launch": (rd + wr) / len(ids), "dram_read_bytes": rd,
           "dram_write_bytes": wr, "ncu_time_us_sum": us,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d GEMM launches of one eager training "
                     "step (%s), divided by the launch count" % (len(ids), os.path.basename(path))}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    json.dump(out, open(os.path.join(root, "profiles", "gemm_dram_traffic.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 147)
