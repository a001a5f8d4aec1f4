#!/usr/bin/env python
"""Device time of `fasta add barcode` on 1 M and 8 M reads (bench.run_add_barcode_leg), one JSON object per size."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

peak = bench.peaks()[0]
for n in (1_000_000, 8_000_000):
    print(json.dumps(bench.run_add_barcode_leg(0, peak, n=n)))
