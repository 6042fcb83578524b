"""Static evidence from the shipped library (no GPU needed): for every kernel of zero_b200/libzero_b200.so its
registers / spill stack / static shared memory (cuobjdump --dump-resource-usage) and how many tensor-core, TMA,
TMEM, barrier and multicast instructions its SASS holds (cuobjdump -sass; mnemonics from B200_PROFILING.md).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zero_b200", "libzero_b200.so")
# tcgen05.mma, TMA loads / stores / reduce-adds, tcgen05.ld / st, tcgen05.commit + mbarrier, mma.sync, cp.async,
# multimem.ld_reduce / multimem.st
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDGMC", "STGMC",
        "REDG"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def short(name, width=110):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)          # drop the argument list
    return name if len(name) <= width else name[:width - 1] + "…"


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(KEYS, 0)
            counts[cur]["insts"] = 0
            continue
        if cur and "/*" in line:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            counts[cur]["insts"] += 1
            op = m.group(1).split(".")[0]
            if op in counts[cur]:
                counts[cur][op] += 1
    names = sorted(counts)
    pretty = demangle(names)
    print("kernels: %d   library: zero_b200/libzero_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)"
          % len(names))
    print("%-112s %4s %5s %6s %6s | %s" % ("kernel", "reg", "stack", "ssmem", "insts", " ".join("%s" % k for k in KEYS)))
    rows = []
    for n in names:
        u, c = usage.get(n, {}), counts[n]
        rows.append((short(pretty[n]), u.get("REG", -1), u.get("STACK", -1), u.get("SHARED", -1), c["insts"],
                     [c[k] for k in KEYS]))
    for name, reg, stack, smem, insts, ks in sorted(rows):
        print("%-112s %4d %5d %6d %6d | %s" % (name, reg, stack, smem, insts,
                                               " ".join(("%%%dd" % len(k)) % v for k, v in zip(KEYS, ks))))
    tc = sum(1 for r in rows if r[5][0] > 0)
    tma = sum(1 for r in rows if r[5][1] > 0)
    mma = sum(1 for r in rows if r[5][8] > 0)
    spill = sorted(r[0] for r in rows if r[2] > 0)
    print("\nkernels with tcgen05.mma (UTCHMMA): %d; with TMA loads (UTMALDG): %d; with mma.sync (HMMA): %d" % (tc, tma, mma))
    print("kernels with a spill stack: %d" % len(spill))
    for s in spill:
        print("   ", s)


if __name__ == "__main__":
    sys.exit(main())
