#!/usr/bin/env python
"""Instruction / stall-sample share per function and phase of the lean engine (reads the output of
tools/ncu_lines.py on stdin: run it with a large top_n)."""
import collections
import re
import sys

import os


def ranges(path):
    """(first line, name) of every function and of every phase marker of the kernel, in file order."""
    out = []
    try:
        lines = open(path).read().splitlines()
    except OSError:
        return out
    for i, ln in enumerate(lines, 1):
        m = re.match(r"(?:static )?__device__ .*?(\w+)\(", ln) or re.match(r"__global__ .*?(\w+)\(", ln)
        if m:
            out.append((i, m.group(1)))
            continue
        m = re.match(r"\s*// ---- (P\d+\w*) ([\w ,+-]+)", ln)
        if m:
            out.append((i, m.group(1) + " " + m.group(2).strip()[:28]))
    return out


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABS = {"sk_fast.cu": ranges(os.path.join(ROOT, "seqkit_b200/csrc/sk_fast.cu")),
        "sk_device.cuh": ranges(os.path.join(ROOT, "seqkit_b200/csrc/sk_device.cuh")),
        "sk_kernels.cu": ranges(os.path.join(ROOT, "seqkit_b200/csrc/sk_kernels.cu"))}


def bucket(f, n):
    name = None
    for lo, nm in TABS.get(f, []):
        if lo <= n:
            name = nm
        else:
            break
    return name or f


kernel = None
agg = collections.OrderedDict()
for ln in sys.stdin:
    if ln.startswith("== "):
        kernel = ln.strip()
        agg[kernel] = collections.defaultdict(lambda: [0.0, 0.0])
        continue
    m = re.match(r"\s+([\d.]+)% samp\s+([\d.]+)% inst.*?(\S+):(\d+)  ", ln)
    if m and kernel:
        b = agg[kernel][bucket(m.group(3), int(m.group(4)))]
        b[0] += float(m.group(1))
        b[1] += float(m.group(2))
    elif kernel and "warp instructions" in ln:
        print(kernel)
        print(ln.rstrip())
for k, d in agg.items():
    print(k)
    for name, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print("   %-18s inst %5.1f%%   samples %5.1f%%" % (name, v[1], v[0]))
