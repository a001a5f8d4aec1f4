# Short GPU call while iterating on one operator: the parity tests that touch it, then the per-launch times
# of a 5-step bench.  usage: bash tools/quick_check.sh "<pytest -k expression>"
timeout 240 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -8
timeout 90 python bench.py --steps 5 --warmup 3 --skip-e2e --skip-cpu --cli-pairs 0 2>/dev/null > gpurun_out/quick_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/quick_bench.json"))
r = d["roofline"]
print("DEMUX1 %.3f ms  DEMUX2 %.3f ms  frac %.3f" % (r["ms_per_launch"], r["other_kernel"]["ms_per_launch"], r["frac"]))
for k, v in r["other_ops"].items():
    print("%s %.3f ms  frac %.3f" % (k, v["ms_per_launch"], v["frac"]))
PY
