"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python scripts/launch_summary.py launches.csv [bench_line.json] > profiles/..._summary.txt"""
import collections
import csv
import json
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rd:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = r[ix["Kernel Name"]].split("(")[0][:60]
    tot[name] += ms; cnt[name] += 1
total = sum(tot.values())
print("# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (gpu__time_duration.sum, --clock-control none)")
print(f"# {sum(cnt.values())} launches, {total:.1f} ms of kernel time (cold-cache, serialised: compare SHARES with the live event timings)")
for k, v in tot.most_common():
    print(f"{k:62s} n={cnt[k]:4d} {v:9.3f} ms {100 * v / total:5.1f}%")
if len(sys.argv) > 2:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r, p = d["roofline"], d["roofline_prepass"]
    print(f"\n# live (CUDA events, same command, no profiler): chain kernel {r['launch_ms']:.2f} ms/step = {100 * r['share_of_step']:.1f}% of the step; "
          f"pool (fill + GEMM) {p['launch_ms']:.2f} ms/step = {100 * p['launch_ms'] / d['ms_per_step']:.1f}%; gpu_launches {d['gpu_launches']}")
