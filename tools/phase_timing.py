"""Per-phase cycle breakdown of the chunk engine (needs the -DSK_PHASE_TIMING build):
   make seqkit_b200/libseqkit_b200_timing.so && SK_LIB=seqkit_b200/libseqkit_b200_timing.so python tools/phase_timing.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from seqkit_b200 import Engine, _lib as L  # noqa: E402

FAST_NAMES = ["ticket", "load", "scan", "lines+lookback", "plan", "layout", "reserve", "assemble", "tables+store"]
NAMES = FAST_NAMES if not os.environ.get("SK_NO_FAST") else ["ticket", "load", "scan", "lines+lookback", "planA(d1)", "slowmatch(d1)", "plan", "layout", "reserve", "assemble",
         "store+tables"]
P = int(os.environ.get("PAIRS", "1000000"))
bcs = bench.make_sheet()
eng = Engine(max_stream_bytes=P * 410 + (1 << 20), max_records=P, max_samples=384, aux_streams=False)
eng.set_sheet(bcs)
n1, n2 = bench.synth_pair(eng, P, 0)
opts = L.DemuxOpts(20, 0, 0, 0, 0)
for it in range(3):
    assert eng.lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
    res = eng.wait()
eng.lib.sk_set_profiling(eng.ctx, 1)
assert eng.lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
res = eng.wait()
print("pairs", P, "pass_ms", list(res.pass_ms)[:2], "cfg", os.environ.get("SK_CFG"))
for which, name in ((0, "DEMUX1"), (1, "DEMUX2")):
    out = (C.c_uint64 * 16)()
    eng.lib.sk_debug_phase_cycles(eng.ctx, 0, which, out)
    tot = sum(out)
    print(name, "total cycles per timing thread-sum:", tot)
    tot = sum(out[:12])
    for i, nm in enumerate(NAMES):
        if out[i]:
            print("   %-16s %5.1f%%" % (nm, 100.0 * out[i] / tot))
    for i, nm in ((12, "within plan: first trim warp done"), (13, "within plan: first header warp done"),
                  (14, "within plan: look-back collected")):
        if out[i]:
            print("   %-36s %5.1f%% of the chunk time" % (nm, 100.0 * out[i] / tot))
