#!/usr/bin/env python
"""Static evidence for the kernels of gimic_b200/libgimic_b200.so as built (CPU only, no GPU needed):
per kernel the ptxas resource line (registers, spills, shared memory; from the *.ptxas.log files the Makefile keeps) and the count of
the SASS mnemonics that matter for the design -- FP64 tensor-core MMAs (DMMA), FP64 FMAs (DFMA), bulk-TMA copies (UBLKCP), mbarrier
traffic (SYNCS), cp.async gathers (LDGSTS), register re-allocation (USETMAXREG), MUFU (exp), local-memory spills (STL/LDL).

    python tools/sass_static.py > profiles/rNN_sass_static.txt

It says what the compiler emitted, not how fast it runs; the measured side is the ncu summaries beside it."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gimic_b200", "libgimic_b200.so")
MNEMONICS = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "LDG", "STG", "LDS", "STS", "STL", "LDL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(.*", "", name)                                   # drop the argument list
    return re.sub(r"^void ", "", name)


def ptxas_table():
    """{mangled kernel: resource line} from the build logs"""
    tab = {}
    for log in sorted(glob.glob(os.path.join(ROOT, "gimic_b200", "csrc", "*.ptxas.log"))):
        cur = None
        spill = ""
        for line in open(log, errors="replace"):
            m = re.search(r"Compiling entry function '([^']+)' for 'sm_100a'", line)
            if m:
                cur = m.group(1); spill = ""
                continue
            if cur and "bytes stack frame" in line:
                spill = line.split("info    :")[-1].strip() if "info" in line else line.strip()
            m = re.search(r"Used (\d+) registers.*", line)
            if m and cur:
                tab[cur] = (m.group(0).strip(), spill.strip())
                cur = None
    return tab


def sass_counts():
    """{mangled kernel: Counter of mnemonics}"""
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for key in MNEMONICS:
                if op == key or (key in ("DMMA", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "BAR") and op.startswith(key)):
                    counts[cur][key] += 1
    return counts


def main():
    if not os.path.exists(SO):
        sys.exit(f"{SO} is missing: python __graft_entry__.py builds it")
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    nvcc = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().split("\n")[-2]
    print(f"# static SASS / ptxas record of gimic_b200/libgimic_b200.so at {head} ({nvcc}; -gencode arch=compute_100a,code=sm_100a)")
    print("# tools/sass_static.py -- what the compiler emitted (CPU container, no GPU involved)\n")
    res, cnt = ptxas_table(), sass_counts()
    names = demangle(sorted(cnt))
    own = [k for k in sorted(cnt) if "cub" not in names[k]]
    for k in own:
        c = cnt[k]
        print(short(names[k]))
        if k in res:
            print("    ptxas: " + res[k][0] + (" | " + res[k][1] if res[k][1] else ""))
        print("    SASS : " + f"{c['total']} instructions; " + ", ".join(f"{m} {c[m]}" for m in MNEMONICS if c[m]))
    print(f"\n# library kernels in the same file (cub radix sort of the Morton keys): {len(cnt) - len(own)}")


if __name__ == "__main__":
    main()
