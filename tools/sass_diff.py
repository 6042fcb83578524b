"""Per-kernel SASS comparison of the CUDA sources between a git revision and the working tree (no GPU needed).

    python tools/sass_diff.py f94db3f              # every csrc/*.cu
    python tools/sass_diff.py f94db3f addln beam   # a few translation units

Used to show that work done without a GPU left the kernels of the last GPU-validated build byte-for-byte alone:
"identical" counts kernels whose instruction streams match, "added" the new (opt-in) ones; any "CHANGED" line is a
kernel of the old revision that compiles differently now and has to be re-validated.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-w"]


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    table, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = []
        elif cur and "/*" in line:
            text = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line)
            text = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", text).strip()
            if text:
                table[cur].append(text)
    return table


def main():
    rev = sys.argv[1]
    units = sys.argv[2:] or sorted(f[:-3] for f in os.listdir(os.path.join(ROOT, "zero_b200", "csrc")) if f.endswith(".cu"))
    changed = 0
    with tempfile.TemporaryDirectory() as tmp:
        for sub in ("zero_b200/csrc", "include"):
            os.makedirs(os.path.join(tmp, sub))
            listing = subprocess.run(["git", "ls-tree", "--name-only", rev, sub + "/"], cwd=ROOT, capture_output=True,
                                     text=True, check=True).stdout.split()
            for path in listing:
                with open(os.path.join(tmp, path), "w") as f:
                    f.write(subprocess.run(["git", "show", "%s:%s" % (rev, path)], cwd=ROOT, capture_output=True,
                                           text=True, check=True).stdout)
        for unit in units:
            old_src = os.path.join(tmp, "zero_b200", "csrc", unit + ".cu")
            if not os.path.exists(old_src):
                print("%-20s new translation unit" % unit)
                continue
            objs = []
            for src, name in ((old_src, "old"), (os.path.join(ROOT, "zero_b200", "csrc", unit + ".cu"), "new")):
                obj = os.path.join(tmp, "%s_%s.o" % (unit, name))
                subprocess.run(["nvcc", *FLAGS, "-c", src, "-o", obj], check=True)
                objs.append(kernels(obj))
            old, new = objs
            diff = [k for k in old if k in new and old[k] != new[k]]
            same = sum(1 for k in old if k in new and old[k] == new[k])
            # a kernel whose template signature grew a defaulted parameter has a new mangled name: match it by its
            # instruction stream
            streams = {}
            for k in new:
                if k not in old:
                    streams.setdefault("\n".join(new[k]), []).append(k)
            renamed, missing = 0, []
            for k in old:
                if k not in new:
                    hit = streams.get("\n".join(old[k]))
                    if hit:
                        hit.pop()
                        renamed += 1
                    else:
                        missing.append(k)
            added = sum(len(v) for v in streams.values())
            print("%-20s identical %3d (+%d renamed)   changed %3d   removed %3d   added %3d" % (
                unit, same, renamed, len(diff), len(missing), added))
            for k in diff + missing:
                print("    CHANGED / REMOVED", k)
            changed += len(diff) + len(missing)
    sys.exit(1 if changed else 0)


if __name__ == "__main__":
    main()
