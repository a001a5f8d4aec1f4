#!/usr/bin/env python
"""Per-kernel SASS summary of a built library: instruction count and the counts of the mnemonics that
show what the code is made of (TMA bulk copies UBLKCP, mbarrier SYNCS, 256-bit stores STG.E.ENL2.256,
dot products IDP.4A, fused add-min/max VIADDMNMX, three-input min/max VIMNMX3 ...).  Runs without a GPU.

  python tools/sass_summary.py [lib.so ...] > profiles/rN_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

libs = sys.argv[1:] or ["seqkit_b200/libseqkit_b200.so"]
WATCH = ["UBLKCP", "SYNCS", "STG.E.ENL2.256", "STG.E.128", "LDS.128", "LDS", "STS", "IDP.4A", "VIADDMNMX", "VIMNMX3", "VIMNMX",
         "SHF", "LOP3", "PRMT", "IADD3", "IMAD", "POPC", "FLO", "VOTE", "SHFL", "BAR", "ATOMG", "RED", "BSSY", "BRA", "LDG", "STG",
         "UTCHMMA", "UTCIMMA", "HMMA"]
for lib in libs:
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur = None
    kernels = collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur is not None:
            op = m.group(1)
            cur["_n"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                    cur[w] += 1
    print("==", lib)
    for k, c in kernels.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:90]
        ops = "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])
        print("  %-90s  instr %5d\n      %s" % (name, c["_n"], ops))
