"""One-page summary of a tools/gpu_validate_queue.sh run: `python tools/summarize_queue.py TAG` reads gpurun_out/TAG_*
and prints (and writes gpurun_out/TAG_summary.txt) what decides which switches become defaults: parity groups, the
training bench per switch, the decode A/B, the L2 / multicast probe."""
import glob
import json
import os
import re
import sys


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "queue"
    base = os.path.join("gpurun_out", tag + "_")
    out = []

    def read(name):
        try:
            return open(base + name).read()
        except OSError:
            return ""
    tests = read("tests.log")
    m = re.findall(r"(\d+ passed[^\n]*|\d+ failed[^\n]*)", tests)
    out.append("default suite      : %s" % (m[-1] if m else "no result"))
    for grp, rc, secs in re.findall(r"=== (.+?) rc=(\d+) in (\d+) s", read("tests_unvalidated.log")):
        out.append("opt-in parity      : %-45s %s (%s s)" % (grp, "PASS" if rc == "0" else "FAIL rc=" + rc, secs))
    for name in ("ln16", "ln8"):
        t = read("tests_%s.log" % name)
        if t:
            m = re.findall(r"(\d+ passed[^\n]*|\d+ failed[^\n]*)", t)
            out.append("add+LN bwd %-7s : %s" % (name, m[-1] if m else "no result"))
    base_val = None
    paths = sorted(glob.glob(base + "bench_*.json"), key=lambda q: (not q.endswith("bench_default.json"), q))
    for path in paths:
        try:
            d = json.loads(open(path).read().strip().splitlines()[-1])
        except Exception:
            out.append("training bench     : %-28s unreadable" % os.path.basename(path))
            continue
        name = os.path.basename(path)[len(tag) + 7:-5]
        if name == "default":
            base_val = d["value"]
        rel = "" if not base_val else "  (%+.1f %% vs default)" % (100.0 * (d["value"] / base_val - 1.0))
        out.append("training bench     : %-28s %10.0f tok/s  %.3f ms/step  GEMM frac %.3f%s" % (
            name, d["value"], d["ms_per_step"], d.get("roofline", {}).get("frac", float("nan")), rel))
    for line in read("decode_ab.jsonl").splitlines():
        try:
            d = json.loads(line)
            out.append("decode A/B         : %-60s %8.0f tok/s  %.3f ms/step" % (d["switches"], d["value"], d["ms_per_step"]))
        except Exception:
            pass
    probe = read("tma_l2_probe.log")
    if probe:
        out.append("L2 / multicast probe:")
        out += ["    " + l for l in probe.splitlines() if l[:1] in "012m" or l.startswith("SMs")]
    text = "\n".join(out)
    print(text)
    with open(base + "summary.txt", "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
