#!/usr/bin/env python
"""Where the `fasta demultiplex` process spends its wall clock (SK_TIMING=1) on a larger sample than the bench's
files+gzip leg: python tools/cli_profile.py [pairs]  -> one line per setting on stdout."""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
bcs = bench.make_sheet()
r1, r2 = bench.host_pairs(bcs, n, seed=11)
top = tempfile.mkdtemp(prefix="skcli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    for name, data in (("sheet.tsv", bench.sheet_text(bcs)), ("r1.fq", r1), ("r2.fq", r2)):
        with open(os.path.join(top, name), "wb") as f:
            f.write(data)
    print("cores", os.cpu_count(), "pairs", n, "bytes", len(r1) + len(r2))
    for env in ({}, {"SK_GZIP_LEVEL": "1"}, {"SK_GZIP_LEVEL": "6"}, {"SK_BATCH_MB": "64"}, {"SK_BATCH_MB": "128", "SK_GZIP_LEVEL": "1"},
                {"SK_NO_COMPACT": "1"}, {"SK_GZIP": "child"}):
        d = os.path.join(top, "out")
        shutil.rmtree(d, ignore_errors=True)
        os.mkdir(d)
        e = dict(os.environ, SK_TIMING="1", **env)
        t0 = time.perf_counter()
        p = subprocess.run([os.path.join(ROOT, "seqkit_b200", "fasta"), "demultiplex", "--trim-by-quality=20", "../sheet.tsv",
                            "../r1.fq", "../r2.fq"], cwd=d, env=e, capture_output=True, timeout=900)
        dt = time.perf_counter() - t0
        gz = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
        tl = [l for l in p.stderr.decode().splitlines() if "timing" in l]
        print(env, "rc", p.returncode, "%.2f s  %.0f k reads/s  gz %.1f MB" % (dt, 2 * n / dt / 1e3, gz / 1e6), tl[-1] if tl else p.stderr[-200:])
finally:
    shutil.rmtree(top, ignore_errors=True)
