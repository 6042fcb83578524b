"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=20):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:52]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print("%-52s n=%4d total %9.1f us avg %8.1f us %5.1f%%" % (k, n, t, t / n, 100 * t / tot))
    print("total %.1f us over %d launches" % (tot, sum(n for n, _ in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20)
