#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch, --set full) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt <warp-level work items in the launch>"""
import collections, csv, io, re, subprocess, sys

rep, out, items = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = srows[1]
ia, isrc = h.index("Instructions Executed"), h.index("Source")
rc = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
cnt, stall, tot = collections.Counter(), collections.Counter(), 0
for r in srows[2:]:
    try:
        n = int(r[ia])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    cnt[m.group(2) if m else "?"] += n
    tot += n
    for i in rc:
        try:
            stall[h[i]] += int(r[i])
        except ValueError:
            pass
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none, one launch; source: %s\n" % rep)
    for k in keys:
        if k in M:
            f.write("%-70s %s %s\n" % (k, M[k][0], M[k][1]))
    f.write("\n# executed warp instructions per warp-level work item (%.0f items in the launch): %.1f total\n" % (items, tot / items))
    fp64 = sum(v for k, v in cnt.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    f.write("# FP64-pipe instructions (DFMA+DMUL+DADD+DSETP) per item: %.1f\n" % (fp64 / items))
    f.write(" ".join("%s %.1f" % (op, n / items) for op, n in cnt.most_common(32)) + "\n")
    st = sum(stall.values())
    f.write("\n# warp stall sampling (all samples)\n" + " ".join("%s %.1f%%" % (k[6:], 100.0 * v / st) for k, v in stall.most_common(12)) + "\n")
print(open(out).read())
