"""Summarise an .ncu-rep (read here on the CPU box): key raw metrics per kernel + SASS hot spots.
usage: python scripts/ncu_summary.py <report.ncu-rep> > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print(f"# ncu summary of {rep}  (--set full --clock-control none; per-launch, cold-cache, serialised)")
for r in rows[2:]:
    print(f"\n## kernel: {r[ix['Kernel Name']][:100]}   (launch id {r[ix['ID']]})")
    for k in KEYS:
        if k in ix and r[ix[k]] != "":
            print(f"  {k:78s} {r[ix[k]]:>18s} {units[ix[k]]}")
    tens = [h for h in hdr if ("tensor" in h or "dmma" in h) and "pct_of_peak_sustained_active" in h and ".avg." in h]
    for h in tens:
        if h not in KEYS and r[ix[h]] not in ("", "0"):
            print(f"  {h:78s} {r[ix[h]]:>18s} {units[ix[h]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page concatenates kernels: split on the "Kernel Name" marker rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(src)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["data"].append(row)
for b in blocks:
    h = {x: i for i, x in enumerate(b["hdr"])}
    data = b["data"]
    ts = sum(int(r[h["# Samples"]]) for r in data) or 1
    ti = sum(int(r[h["Instructions Executed"]]) for r in data) or 1
    print(f"\n## SASS profile: {b['name'][:100]}\n  SASS instructions {len(data)}, warp-level instructions executed {ti}, samples {ts}")
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        t = r[h["Source"]].split()
        o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        op[o] += int(r[h["Instructions Executed"]]); ops[o] += int(r[h["# Samples"]])
    print("  opcode mix (share of executed / share of stall samples):")
    for o, c in op.most_common(16):
        print(f"    {o:10s} {c / ti * 100:5.1f}%  {ops[o] / ts * 100:5.1f}%")
    st = [x for x in b["hdr"] if x.startswith("stall_") and "Not Issued" not in x]
    agg = {x: sum(int(r[h[x]]) for r in data) for x in st}
    print("  stall reasons (% of samples): " + ", ".join(f"{k[6:]} {v / ts * 100:.1f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
