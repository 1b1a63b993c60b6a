#!/usr/bin/env python
"""Static evidence for profiles/: per kernel instance of libprosail_b200.so
  (1) registers / spill bytes / shared memory from `ptxas -v` (the library is recompiled to /tmp with -Xptxas -v), and
  (2) an opcode histogram of the shipped SASS (`cuobjdump -sass`): FP64-pipe instructions (DFMA / DMUL / DADD / DSETP),
      three-register FP64 instructions (which take a third pipe cycle), SFU (MUFU, F2F, I2F, F2I), shared
      / local / global memory instructions, and the Blackwell-specific mnemonics (UBLKCP = cp.async.bulk, SYNCS = mbarrier,
      FFMA2 / FMUL2 / FADD2 = packed fp32, UTC*MMA / LDTM / UTMALDG = tcgen05 / TMEM / TMA tensor copies).
Whole-function counts (rare paths included), so they bound -- not equal -- the executed counts of the ncu summaries.
usage: python tools/sass_report.py > profiles/r2_sass_report.txt      (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
LIB = os.path.join(ROOT, "pyrism_b200", "lib", "libprosail_b200.so")


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        if len(out) == len(names):
            return out
    except Exception:
        pass
    return names


def short(d):
    d = d.replace("void pb::", "").replace("pb::", "").replace("(unsigned int)", "").replace("(bool)", "").replace("(int)", "")
    return d.split(">(")[0] + (">" if ">(" in d else "")


def ptxas_table():
    out = subprocess.run(["make", "ptxas"], capture_output=True, text=True, cwd=ROOT)
    txt = out.stdout + out.stderr
    rows, cur = [], None
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
            cur["stack"], cur["st"], cur["ld"] = int(m.group(1)), int(m.group(2)), int(m.group(3))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            m2 = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(m2.group(1)) if m2 else 0
    return rows


def sass_histograms():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), {"ops": collections.Counter(), "three": 0, "n": 0})
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P[0-9T]\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)\s*(.*?);", line)
        if not m:
            continue
        op, mods, operands = m.group(2), m.group(3), m.group(4)
        key = op
        if op == "MUFU" and mods:
            key = op + "." + mods.split(".")[1]
        cur["ops"][key] += 1
        cur["n"] += 1
        if op in ("DFMA", "DMUL", "DADD"):
            regs = set()
            for o in operands.split(",")[1:]:
                o = o.strip().lstrip("-|").rstrip("|")
                if re.match(r"R\d+", o):
                    regs.add(o.split(".")[0])
            if len(regs) >= 3:
                cur["three"] += 1
    return funcs


def main():
    rows = ptxas_table()
    dem = demangle([r["name"] for r in rows])
    print("# ptxas -v (sm_100a), one line per kernel entry: registers, stack frame, spill stores / loads (bytes), static shared memory")
    for r, d in sorted(zip(rows, dem), key=lambda x: short(x[1])):
        print("%-74s regs %3d  stack %4d  spill st/ld %4d/%4d  smem %6d" % (short(d)[:74], r.get("regs", -1), r.get("stack", 0), r.get("st", 0),
                                                                         r.get("ld", 0), r.get("smem", 0)))
    print()
    print("# cuobjdump -sass of the shipped library: whole-function opcode counts per kernel instance")
    funcs = sass_histograms()
    dem = demangle(list(funcs.keys()))
    tot = collections.Counter()
    for (name, f), d in sorted(zip(funcs.items(), dem), key=lambda x: short(x[1])):
        o = f["ops"]
        fp64 = o["DFMA"] + o["DMUL"] + o["DADD"] + o["DSETP"]
        mufu = sum(v for k, v in o.items() if k.startswith("MUFU"))
        cvt = o["F2F"] + o["I2F"] + o["F2I"]
        packed = o["FFMA2"] + o["FMUL2"] + o["FADD2"]
        tc = sum(v for k, v in o.items() if k.startswith("UTC") or k in ("LDTM", "STTM", "UTMALDG", "UTMASTG"))
        print("%-74s total %5d | DFMA %4d DMUL %4d DADD %4d DSETP %3d (FP64 %4d, three-register %3d) | MUFU %3d F2F/I2F/F2I %3d | FFMA %4d packed-f32 %4d | "
              "LDS %3d STS %3d LDL %3d STL %3d LDG %3d STG %3d | UBLKCP %d SYNCS %d | tensor (UTC*MMA/LDTM/UTMA*) %d"
              % (short(d)[:74], f["n"], o["DFMA"], o["DMUL"], o["DADD"], o["DSETP"], fp64, f["three"], mufu, cvt, o["FFMA"], packed, o["LDS"], o["STS"], o["LDL"],
                 o["STL"], o["LDG"], o["STG"], o["UBLKCP"], o["SYNCS"], tc))
        tot.update(o)
    print()
    print("# library totals: " + " ".join("%s %d" % kv for kv in tot.most_common(40)))


if __name__ == "__main__":
    sys.exit(main())
