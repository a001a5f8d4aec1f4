#!/usr/bin/env python
"""Per-source-line summary of an ncu report (run here, no GPU needed).

ncu's CSV export of the source page carries per-SASS-instruction metrics but no line numbers, so the
kernel's SASS is re-disassembled from the built library with `nvdisasm -g` (needs -lineinfo at
compile time) and joined by instruction order.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [seqkit_b200/libseqkit_b200.so] [top_n]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "seqkit_b200/libseqkit_b200.so"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sort_by = 1 if (len(sys.argv) > 4 and sys.argv[4] == "inst") else 0  # "inst": order by executed instructions

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
funcs = {}  # mangled name -> list of (line, inlined_from_line)
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, line = None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            line = None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            # keep the outermost sk_kernels.cu line when the instruction comes from an inlined header
            f, n, rest = m.group(1), int(m.group(2)), m.group(3)
            m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
            cand = [(f, n)] + [(a, int(b)) for a, b in m2]
            own = [c for c in cand if c[0].endswith(".cu")]
            line = (own[0] if own else cand[0])
            continue
        if cur is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            cur.append(line)

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernels, hdr = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels.append({"name": r[1], "rows": []})
        hdr = None
    elif kernels and hdr is None:
        hdr = r
        kernels[-1]["hdr"] = r
    elif kernels:
        kernels[-1]["rows"].append(r)


def demangle_match(kname, funcs):
    # match by template arguments: Cfg name and OP number
    m = re.search(r"sk_warp_kernel<\(int\)(\d+), \(int\)(\d+)>", kname)
    if m:
        for f in funcs:
            if "sk_warp_kernel" in f and ("Li%sELi%sE" % (m.group(1), m.group(2))) in f:
                return f
        return None
    m = re.search(r"sk_chunk_kernel<sk::(\w+), \(int\)(\d+), (unsigned int|unsigned long)", kname)
    if not m:
        return None
    cfg, op, wt = m.group(1), m.group(2), "j" if m.group(3) == "unsigned int" else "m"
    for f in funcs:
        if "sk_chunk_kernel" in f and ("%d%s" % (len(cfg), cfg)) in f and ("Li%sE%sE" % (op, wt)) in f:
            return f
    return None


src_cache = {}


def src_line(f, n):
    if f not in src_cache:
        try:
            src_cache[f] = open(f).read().splitlines()
        except OSError:
            src_cache[f] = []
    s = src_cache[f]
    return s[n - 1].strip() if 0 < n <= len(s) else ""


seen = set()
for K in kernels:
    if K["name"] in seen:
        continue
    seen.add(K["name"])
    h = K["hdr"]
    f = demangle_match(K["name"], funcs)
    print("==", K["name"])
    if f is None or len(funcs[f]) != len(K["rows"]):
        print("   cannot join with nvdisasm (%s, %d vs %d instructions)" % (f, len(funcs.get(f, [])), len(K["rows"])))
        continue
    col = {n: h.index(n) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_barrier",
                                   "stall_short_sb", "stall_long_sb", "stall_wait", "L1 Wavefronts Shared Excessive")}
    agg = collections.defaultdict(lambda: [0.0] * 8)
    for r, ln in zip(K["rows"], funcs[f]):
        a = agg[ln]
        for k, n in enumerate(col):
            try:
                a[k] += float(r[col[n]] or 0)
            except ValueError:
                pass
    tot = [sum(a[k] for a in agg.values()) or 1.0 for k in range(8)]
    print("   warp instructions %.0f, samples %.0f, barrier samples %.0f (%.0f%%), lane efficiency %.1f/32"
          % (tot[1], tot[0], tot[3], 100 * tot[3] / tot[0], tot[2] / tot[1]))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][sort_by])[:topn]:
        fn, n = ln if ln else ("?", 0)
        print("   %5.1f%% samp %5.1f%% inst  bar %5.1f%% ssb %4.1f%% wait %4.1f%% xsw %6.0fk  %s:%d  %s"
              % (100 * a[0] / tot[0], 100 * a[1] / tot[1], 100 * a[3] / tot[0], 100 * a[4] / tot[0], 100 * a[6] / tot[0],
                 a[7] / 1e3, os.path.basename(fn), n, src_line(fn, n)[:90]))
