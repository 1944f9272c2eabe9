#!/usr/bin/env python
"""Static evidence of what the shipped library runs on: per kernel, the SASS mnemonics that prove tcgen05 / tensor
memory / TMA use (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA
load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier), the special-function and FMA counts, and the resource usage
recorded in the cubin (registers, static shared memory, local-memory stack = spills).  No GPU needed:

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unseenobjectclustering_b200", "libuoc_b200.so")

COLUMNS = [("UTCHMMA", r"\bUTC[HQI]?MMA"), (".2CTA", r"\bUTC[HQI]?MMA\S*\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
           ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG|\bUTMAREDG"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"),
           ("MUFU", r"\bMUFU"), ("FFMA", r"\bFFMA"), ("HMMA/IMMA", r"\b[HI]MMA\.|\bHMMA\b")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("uoc::", "")
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)                      # drop the argument list
    return name


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            counts[cur]["_instr"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        counts[cur]["_instr"] += 1
        for col, pat in COLUMNS:
            if re.search(pat, ins):
                counts[cur][col] += 1
    usage = {}
    fn = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+?):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            usage[fn] = tuple(int(x) for x in m.groups())
            fn = None
    names = demangle(list(counts))
    print("# SASS summary of `unseenobjectclustering_b200/libuoc_b200.so` (%s), `tools/sass_summary.py`\n" % ", ".join(arch))
    print("Counts of instructions per kernel in the shipped cubin (`cuobjdump -sass`), resources from `cuobjdump -res-usage`")
    print("(REG = registers per thread, SMEM = static shared memory in bytes -- the tcgen05 kernels take their tiles from")
    print("dynamic shared memory, STACK = bytes of local-memory stack per thread, blank = none: no spills, no local arrays).  Template instances are listed")
    print("separately.  UTCHMMA = `tcgen05.mma` (kind::f16, bf16 operands), `.2CTA` = `cta_group::2`, LDTM / STTM =")
    print("`tcgen05.ld` / `tcgen05.st`, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier.\n")
    hdr = ["kernel", "instr"] + [c for c, _ in COLUMNS] + ["REG", "SMEM", "STACK"]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    tot = collections.Counter()
    rows = []
    for fn, c in counts.items():
        u = usage.get(fn, (0, 0, 0, 0))
        rows.append((short(names[fn]), c, u))
    rows.sort(key=lambda r: (-(r[1]["UTCHMMA"] > 0), -r[1]["UTMALDG"], r[0]))
    for name, c, u in rows:
        cells = [name if len(name) < 90 else name[:87] + "...", str(c["_instr"])] + [str(c[col]) if c[col] else "" for col, _ in COLUMNS]
        cells += [str(u[0]), str(u[2]), str(u[1]) if u[1] else ""]
        print("| " + " | ".join(cells) + " |")
        tot.update(c)
    print("\nTotals over %d kernels: " % len(rows) + ", ".join("%s %d" % (col, tot[col]) for col, _ in COLUMNS) + ".")
    deps = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    libs = sorted(set(re.findall(r"^\s*(\S+) =>", deps, re.M)))
    print("\n`ldd`: " + ", ".join(libs) + " -- no cuBLAS / cuDNN / NCCL / torch in the product library.")


if __name__ == "__main__":
    main()
