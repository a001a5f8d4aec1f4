#!/usr/bin/env python
"""Device time of the single-stream operators on resident synthetic reads: trim / mask by quality (8 M reads) and
add barcode (1 M and 8 M reads), CUDA events around every pass.  One JSON object per line.
  SK_LIB=seqkit_b200/variants/x.so SK_TRIM_GATHER=0 python tools/stream_time.py [reads]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from seqkit_b200 import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
peak = bench.peaks()[0]
with Engine(max_stream_bytes=n * 420 + (1 << 20), max_records=n, max_samples=8, aux_streams=False) as eng:
    lib = eng.lib
    n_in = eng.synth(0, n, seed=7, mate=1, with_bc=False)
    lib.sk_set_profiling(eng.ctx, 1)
    for name, fn in (("trim", lib.sk_trim_by_quality), ("mask", lib.sk_mask_by_quality)):
        ms, reps = 0.0, 5
        for i in range(reps + 2):
            assert fn(eng.ctx, 0, 20, 0) == 0
            r = eng.wait()
            assert r.status == 0 and r.n_records == n
            if i >= 2:
                ms += r.pass_ms[0] / reps
        b = n_in + int(r.out_bytes[0])
        print(json.dumps({"op": name, "reads": n, "ms": round(ms, 4), "frac": round(b / (ms * 1e-3) / 1e9 / peak, 4),
                          "launches": int(r.gpu_launches), "engine_bits": int(r.reserved),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("SK_")}}))
if "--no-addbc" not in sys.argv:
    for m in (1_000_000, n):
        print(json.dumps(bench.run_add_barcode_leg(0, peak, n=m)))
