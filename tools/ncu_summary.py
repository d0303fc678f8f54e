#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) for profiles/: key metrics, stall reasons, hottest SASS lines."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2 + kidx]
def g(name):
    return r[hdr.index(name)] if name in hdr else "n/a"
print("kernel:", g("Kernel Name"), "grid", g("Grid Size"), "block", g("Block Size"))
for m in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "sm__cycles_active.avg", "smsp__inst_executed.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"]:
    if m in hdr:
        print(f"  {m:75s} {g(m):>16s} {units[hdr.index(m)]}")
st = []
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        try: st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except: pass
print("stall reasons (warps per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
si, ii = h2.index("Warp Stall Sampling (All Samples)"), h2.index("Instructions Executed")
data = []
for idx, rr in enumerate(rows[2:]):
    if len(rr) > max(si, ii):
        try: data.append((int(rr[si] or 0), int(rr[ii] or 0), idx, rr[1].strip()))
        except: pass
tot = sum(d[0] for d in data)
print(f"source page: {len(data)} SASS instructions, {tot} stall samples; hottest:")
for s_, i_, idx, t in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"  {100*s_/tot:5.1f}%  exec={i_:8d}  #{idx:4d}  {t[:100]}")
with open("/tmp/sass_samples.txt", "w") as f:
    f.write("\n".join(f"{idx:5d} {s_:6d} {i_:8d} {t}" for s_, i_, idx, t in sorted(data, key=lambda d: d[2])))
