#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store), UTCBAR (tcgen05.commit), UTMAPF (tensormap prefetch),
RED (fp32 vector reductions of split-K wgrad), plus registers per thread.  Runs on the build host (no GPU needed):

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mipnerf360_b200", "lib", "libmip360_b200.so")
MNEMONICS = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "UTCATOMSWS", "SYNCS", "REDG", "MUFU", "HMMA")


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[fn]["_all"] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k + "."):
                    counts[fn][k] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a), {len(counts)} kernels")
    print(f"# {'kernel':100s} {'instr':>6s} {'regs':>4s} " + " ".join(f"{k:>8s}" for k in MNEMONICS))
    tot = collections.Counter()
    for fn, c in sorted(counts.items(), key=lambda kv: demangle(kv[0])):
        name = re.sub(r"\(.*", "", demangle(fn)).replace("mip360::", "")
        print(f"{name[:102]:102s} {c['_all']:6d} {regs.get(fn, 0):4d} " + " ".join(f"{c[k]:8d}" for k in MNEMONICS))
        tot.update(c)
    print(f"{'TOTAL':102s} {tot['_all']:6d} {'':4s} " + " ".join(f"{tot[k]:8d}" for k in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())
