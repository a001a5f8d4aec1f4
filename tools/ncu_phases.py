#!/usr/bin/env python
"""Per-phase summary (instructions, samples, stall reasons) of the warp-engine kernels in an ncu report.
  python tools/ncu_phases.py gpurun_out/prof.ncu-rep [lib]
Joins ncu's source page with `nvdisasm -g` line info like tools/ncu_lines.py, then buckets every SASS
instruction by the innermost source location (file:line -> phase table below)."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "seqkit_b200/libseqkit_b200.so"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
funcs = {}
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    cur, line = None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            line = None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            f, n, rest = m.group(1), int(m.group(2)), m.group(3)
            m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
            chain = [(os.path.basename(f), n)] + [(os.path.basename(a), int(b)) for a, b in m2]
            line = chain
            continue
        if cur is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            cur.append(line)

# markers in sk_warp.cu: a phase starts at the line holding the marker text
src = open("seqkit_b200/csrc/sk_warp.cu").read().splitlines()
marks = [("setup", "template <int OP, int NWMAX>"), ("load", "// ---- load the window"), ("scan", "// ---- newline scan"),
         ("count+lines", "uint32_t cnt_all = 0;"), ("framing", "// ---- framing."), ("nrec", "uint32_t nrec = 0;"),
         ("plan:trim", "// ---- plan: one lane per record"), ("plan:header", "// trim / mask by quality: failure kind in errk, output length in slen"),
         ("verify/lookback", "// ---- the guess is verified"), ("outcome", "// ---- outcome of every record"),
         ("layout", "// ---- place of the record"), ("emit:header", "// ---- emit"), ("emit:body", "// body: "),
         ("tables", "if (emit) {\n"), ("tail", "if (wrong) continue;")]
starts = []
for name, text in marks:
    t = text.strip()
    for i, l in enumerate(src):
        if t in l and (not starts or i + 1 > starts[-1][0]):
            starts.append((i + 1, name))
            break
starts.sort()

last_phase = ["?"]
def phase_of(chain):
    ph = phase_of0(chain)
    if ph == "?":
        return last_phase[0]  # deeper inline frames are not in the line table: stay with the surrounding phase
    last_phase[0] = ph
    return ph
def phase_of0(chain):
    if not chain:
        return "?"
    for f, n in chain:  # innermost first; the sk_warp.cu frame decides the phase
        if f == "sk_warp.cu":
            if n < starts[0][0]:
                return "gcopy" if 45 <= n <= 95 else "pre"
            ph = [nm for s, nm in starts if s <= n][-1]
            return ph
    return "?"

def sub_of(chain):
    f, n = chain[0]
    return f

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernels, hdr = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels.append({"name": r[1], "rows": []}); hdr = None
    elif kernels and hdr is None:
        hdr = r; kernels[-1]["hdr"] = r
    elif kernels:
        kernels[-1]["rows"].append(r)
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_lg", "stall_mio", "stall_math", "stall_not_selected", "stall_selected",
          "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_barrier", "stall_membar", "stall_sleep", "stall_misc", "stall_drain", "stall_tex"]
seen = set()
for K in kernels:
    if K["name"] in seen: continue
    seen.add(K["name"])
    last_phase[0] = "?"
    m = re.search(r"sk_warp_kernel<\(int\)(\d+), \(int\)(\d+)>", K["name"])
    if not m:
        continue
    f = [x for x in funcs if "sk_warp_kernel" in x and ("Li%sELi%sE" % (m.group(1), m.group(2))) in x][0]
    if len(funcs[f]) != len(K["rows"]):
        print("cannot join", K["name"]); continue
    h = K["hdr"]
    ci = {n: h.index(n) for n in ["# Samples", "Instructions Executed", "Thread Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"] + STALLS}
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    tot = collections.defaultdict(float)
    for r, chain in zip(K["rows"], funcs[f]):
        ph = phase_of(chain)
        for n, i in ci.items():
            try: v = float(r[i] or 0)
            except ValueError: v = 0.0
            agg[ph][n] += v; tot[n] += v
    print("==", K["name"], " warp-instr %.0f  samples %.0f  lanes/instr %.1f" % (tot["Instructions Executed"], tot["# Samples"], tot["Thread Instructions Executed"] / tot["Instructions Executed"]))
    print("   stalls overall: " + "  ".join("%s %.1f%%" % (s[6:], 100 * tot[s] / tot["# Samples"]) for s in STALLS if tot[s] / tot["# Samples"] > 0.01))
    order = [nm for _, nm in starts] + ["gcopy", "pre", "?"]
    for ph in order:
        a = agg.get(ph)
        if not a: continue
        top = sorted(((a[s], s[6:]) for s in STALLS), reverse=True)[:3]
        print("   %-16s inst %5.1f%%  samp %5.1f%%  lanes %4.1f  smem-wavefronts %4.1fM (ideal %4.1fM)  %s" % (
            ph, 100 * a["Instructions Executed"] / tot["Instructions Executed"], 100 * a["# Samples"] / tot["# Samples"],
            a["Thread Instructions Executed"] / max(a["Instructions Executed"], 1), a["L1 Wavefronts Shared"] / 1e6, a["L1 Wavefronts Shared Ideal"] / 1e6,
            " ".join("%s %.0f%%" % (n, 100 * v / max(a["# Samples"], 1)) for v, n in top)))
