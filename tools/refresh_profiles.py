#!/usr/bin/env python
"""Copies the outputs of tools/round_measure.sh (gpurun_out/) into profiles/: the bench line, the ncu launch
list with a per-kernel summary, the per-phase / per-line summary of the full capture and the DRAM-traffic /
issue / pipe figures of the hot kernels (profiles/<round>_traffic.json, which bench.py reads for roofline.traffic;
<round> = the tag up to its first underscore).  usage: python tools/refresh_profiles.py r2_warp"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_warp"

shutil.copy(os.path.join(G, "rm_bench.json"), os.path.join(P, tag + "_bench.json"))
shutil.copy(os.path.join(G, "rm_launches.csv"), os.path.join(P, tag + "_launches.csv"))

rows = [r for r in csv.reader(open(os.path.join(G, "rm_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
with open(os.path.join(P, tag + "_launch_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 --skip-e2e "
            "--skip-cpu --cli-pairs 0 (profiles/%s_launches.csv)\ncold-cache, serialised per-launch times; what matters is "
            "each kernel's share of a step\n(sk_warp_kernel<4,8> = DEMUX1, <5,8> = DEMUX2: the two launches of a bench step; "
            "<1,8> = trim, <2,8> = mask: roofline.other_ops; synth_* = input generation, outside the timed region)\n\n" % tag)
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-72s launches %3d  total %9.1f us  per launch %9.1f us  share %5.1f%%\n" % (k[:70], n, t, t / n, 100 * t / tot))

rep = os.path.join(G, "prof_head.ncu-rep")
with open(os.path.join(P, tag + "_ncu_summary.txt"), "w") as f:
    for tool in ("ncu_phases.py", "ncu_lines.py"):
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), rep], capture_output=True, text=True,
                               cwd=ROOT).stdout)
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = raw[0], raw[1]
c = hdr.index


def val(r, n):
    return float(r[c(n)].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[c(n)], 1.0)


d = collections.OrderedDict(source="ncu --set full --clock-control none, bench.py --pairs 1000000 --steps 1 --warmup 3, one launch "
                            "each (tools/ncu_run.sh; summary in profiles/%s_ncu_summary.txt)" % tag, pairs_per_launch=1000000)
keys = {"dram_bytes_per_launch": None, "gpu_time_us_under_ncu": "gpu__time_duration.sum",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "icc_hit_rate_pct": "sm__icc_request_hit_rate.pct",
        "gcc_instruction_requests_pct_of_peak": "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "warp_instructions": "smsp__inst_executed.sum",
        "pipe_alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "pipe_fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lanes_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio"}
for k in keys:
    d[k] = {}


def take(nm, r):
    for k, m in keys.items():
        try:
            d[k][nm] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") if m is None else val(r, m)
        except Exception:
            d[k][nm] = None


for nm, r in zip(("DEMUX1", "DEMUX2"), raw[2:]):
    take(nm, r)
for nm, fn in (("TRIM", "prof_trim.ncu-rep"), ("MASK", "prof_mask.ncu-rep"), ("COMPACT_MOVE", "prof_move.ncu-rep"),
               ("ADDBC", "prof_addbc.ncu-rep")):
    rp = os.path.join(G, fn)
    if not os.path.exists(rp):
        continue
    rr = list(csv.reader(subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    if len(rr) > 2:
        hdr, units = rr[0], rr[1]
        c = hdr.index
        take(nm, rr[2])
d["note"] = "TRIM = the trim kernel alone (its scan and gather launches are separate kernels); *_pct = percent of the pipe's peak"
json.dump(d, open(os.path.join(P, tag.split("_")[0] + "_traffic.json"), "w"), indent=1)
print(json.dumps(d, indent=1))
