# Full GPU suite, smoke, then the bench line (tools/ output under gpurun_out/).  usage: bash tools/suite_and_bench.sh [bench args]
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/rm_tests.log
cat gpurun_out/rm_tests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py "$@" > gpurun_out/rm_bench.json 2> gpurun_out/rm_bench.err
tail -c 1500 gpurun_out/rm_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/rm_bench.json"))
except Exception as e:
    print("no bench line:", e); raise SystemExit
r = d["roofline"]
print("value %.3g reads/s (with compaction %.3g)  step %.3f ms  D1 %.3f ms D2 %.3f ms compact %.3f ms  frac %.3f" % (
    d["value"], d["value_demux_passes_only"], d["ms_per_step"], r["ms_per_launch"], r["other_kernel"]["ms_per_launch"],
    d["value_breakdown"]["compact_ms"], r["frac"]))
print("e2e", json.dumps(d["e2e"])[:400])
print("files", json.dumps(d["e2e_files_gzip"])[:600])
print("configs", json.dumps(d["configs"])[:2500])
print("verify", d["verify"], "clocks", d["clocks"])
for k, v in r.get("other_ops", {}).items():
    print(k, "%.3f ms frac %.3f" % (v["ms_per_launch"], v["frac"]))
PY
