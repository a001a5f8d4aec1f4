#!/usr/bin/env python
"""The `fasta` binary on every visible GPU against SK_GPUS=1 on the same files (diagnostic for multi-GPU boxes):
exit status, stderr tail, decompressed outputs.  python tools/multi_gpu_cli_check.py"""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzzgen as G  # noqa: E402

sheet, bcs = G.make_sheet(5, 48, 20, umi=8, dual=True)
r1, r2 = G.clean_pairs(77, 9000, bcs, p_sub=0.03, p_random=0.05)
top = tempfile.mkdtemp(prefix="skmg_")
outs = []
for tag, env in (("all", {"SK_GPUS": "64"}), ("one", {"SK_GPUS": "1"})):
    d = os.path.join(top, tag)
    os.mkdir(d)
    for name, data in (("sheet.tsv", sheet), ("r1.fq", r1), ("r2.fq", r2)):
        open(os.path.join(d, name), "wb").write(data)
    e = dict(os.environ, SK_BATCH_MB="1", SK_TIMING="1", **env)
    p = subprocess.run([os.path.join(ROOT, "seqkit_b200", "fasta"), "demultiplex", "sheet.tsv", "r1.fq", "r2.fq"], cwd=d, env=e,
                       capture_output=True, timeout=300)
    files = {f: gzip.decompress(open(os.path.join(d, f), "rb").read()) for f in sorted(os.listdir(d)) if f.endswith(".fq.gz")}
    print(tag, "rc", p.returncode, "files", len(files), "bytes", sum(len(v) for v in files.values()))
    print("  stderr tail:", p.stderr.decode("utf-8", "replace")[-600:].replace("\n", "\n    "))
    outs.append((p.returncode, p.stderr.replace(b"seqkit_b200 timing", b"").split(b"\n")[:3], files))
print("identical files:", outs[0][2] == outs[1][2])
if outs[0][2] != outs[1][2]:
    for k in outs[1][2]:
        a, b = outs[0][2].get(k), outs[1][2][k]
        if a != b:
            print("  first differing file", k, len(a or b""), len(b))
            break
