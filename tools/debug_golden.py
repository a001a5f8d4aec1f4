import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_util as GU
from seqkit_b200 import Engine
from oracle import pyoracle as O
eng = Engine(max_stream_bytes=8 << 20, max_records=1 << 16, max_samples=64)
n = 0
for c in GU.cases():
    if c["op"] not in ("trim", "mask"):
        continue
    data = GU.blob(c["input"])
    fn = eng.trim_by_quality if c["op"] == "trim" else eng.mask_by_quality
    got = fn(data, c["min_baseq"])
    want = (c["exit_code"], GU.blob(c["stdout"]))
    if got[0] != want[0] or got[1] != want[1]:
        n += 1
        if n > 3:
            continue
        print("MISMATCH", c["tag"], c["op"], c["min_baseq"], "exit", got[0], want[0], "len", len(got[1]), len(want[1]))
        a, b = got[1], want[1]
        i = next((k for k in range(min(len(a), len(b))) if a[k] != b[k]), min(len(a), len(b)))
        print("  first diff at", i)
        print("  got :", a[max(0, i - 80):i + 80])
        print("  want:", b[max(0, i - 80):i + 80])
        # which record
        recs_want = b[:i].count(b"\n") // 4
        lines = data.split(b"\n")
        print("  input record", recs_want, lines[4 * recs_want:4 * recs_want + 4])
print("mismatches", n)
