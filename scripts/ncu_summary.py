#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed): key counters, stall mix, hottest SASS lines.
usage: scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_hot_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
print("kernel:", m.get("Kernel Name", ("?",))[0][:110])
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max"]
for k in keys:
    if k in m:
        print(f"  {k:80s} {m[k][0]:>16s} {m[k][1]}")
print("stalls (warps per issue-active):")
st = []
for h, (v, u) in m.items():
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        try:
            st.append((float(v.replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
for v, n in sorted(st, reverse=True)[:9]:
    print(f"  {n:28s} {v:7.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[isamp]), int(r[iex]), r[ia], r[isrc], r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print(f"hot SASS lines (of {tot} samples, {sum(d[1] for d in data)} warp-instructions, {len(data)} static):")
for s_, ex, a, srcl, r in sorted(data, key=lambda d: -d[0])[:nhot]:
    top = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    print(f"  {100 * s_ / tot:5.1f}% ex={ex:9d} {a[-5:]} {srcl[:64]:64s} {[(n, c) for c, n in top if c]}")
chunk = 64
print("by address range (samples %, avg executions per static instruction):")
for i in range(0, len(data), chunk):
    seg = data[i:i + chunk]
    print(f"  {seg[0][2][-5:]} {100 * sum(d[0] for d in seg) / tot:5.1f}%  ex/inst={sum(d[1] for d in seg) / len(seg):10.0f}  {seg[0][3][:48]}")
