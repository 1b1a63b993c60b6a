#!/usr/bin/env python
"""Summarise `make ptxas` output: registers / spills / smem per kernel entry (development tool)."""
import re
import subprocess
import sys

out = subprocess.run(["make", "ptxas"], capture_output=True, text=True)
txt = out.stdout + out.stderr
cur = None
rows = []
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1)}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["spill"] = (int(m.group(2)), int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
try:
    dem = subprocess.run(["cu++filt"] + [r["name"] for r in rows], capture_output=True, text=True).stdout.splitlines()
except Exception:
    dem = [r["name"] for r in rows]
for r, d in zip(rows, dem):
    d = d.replace("void pb::", "").replace("(unsigned int)", "").replace("(bool)", "").replace("(int)", "").split(">(")[0] + ">"
    print("%-60s regs %3d  spill st/ld %4d/%4d" % (d[:60], r.get("regs", -1), r.get("spill", (0, 0))[0], r.get("spill", (0, 0))[1]))
if "error" in txt:
    print(txt[-3000:])
    sys.exit(1)
