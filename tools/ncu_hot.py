"""Top stall sites of the kernels in an .ncu-rep (source page, SASS): python tools/ncu_hot.py REP [kernel-substr] [top]"""
import csv
import io
import subprocess
import sys


def main(rep, sub="", top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {"name": line.split(",", 1)[1][:90], "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    seen = set()
    for b in blocks:
        if sub not in b["name"] or b["name"] in seen:
            continue
        seen.add(b["name"])
        rows = list(csv.DictReader(io.StringIO("\n".join(b["lines"]))))
        tot = sum(int(r["# Samples"] or 0) for r in rows)
        print("==", b["name"], "samples", tot, "instrs", len(rows))
        stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
        agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
        print("  ", ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
        order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
        for i in sorted(order):
            r = rows[i]
            why = sorted(((int(r[k] or 0), k[6:]) for k in stalls), reverse=True)[:2]
            print("  %4d %5.1f%%  %-70s %s" % (i, 100.0 * int(r["# Samples"] or 0) / max(tot, 1), r["Source"].strip()[:70],
                                            " ".join("%s=%d" % (k, v) for v, k in why if v)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 25)
