#!/usr/bin/env python
"""Ablation timings of the demultiplex passes (device time per pass, CUDA events around the kernels):
fused trim on/off, output on/off.  Tells which part of a tile's work the kernel time follows.
  python tools/ablate.py [pairs]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from seqkit_b200 import Engine, _lib as L  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
bcs = B.make_sheet()
eng = Engine(device=0, max_stream_bytes=P * 410 + (1 << 20), max_records=P, n_slots=1, max_samples=B.N_SAMPLES, aux_streams=False)
eng.set_sheet(bcs)
B.synth_pair(eng, P, 0)
lib = eng.lib
lib.sk_set_profiling(eng.ctx, 1)
for name, fused, no_out in (("fused trim + output", B.MIN_BASEQ, 0), ("no trim, output", -1, 0), ("fused trim, no output", B.MIN_BASEQ, 1),
                            ("no trim, no output", -1, 1)):
    opts = L.DemuxOpts(fused, 0, 0, no_out, 0)
    ms = [0.0, 0.0]
    n = 6
    for i in range(n + 2):
        assert lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
        r = eng.wait()
        if i >= 2:
            ms[0] += r.pass_ms[0] / n
            ms[1] += r.pass_ms[1] / n
    print("%-24s DEMUX1 %.3f ms  DEMUX2 %.3f ms  (%d pairs, engine bits %d)" % (name, ms[0], ms[1], P, r.reserved))
eng.close()
