"""Summarise registers / spills per solver_kernel instantiation from the -Xptxas -v build logs.
usage: python tools/ptxas_summary.py [filter-substring]"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = {0: "FWD", 1: "STORE", 2: "REV_S", 3: "REV_GRAD"}
KINDS = {0: "LIN", 1: "RBF", 2: "STATIC", 3: "INC"}
flt = sys.argv[1] if len(sys.argv) > 1 else ""
rows = []
for log in sorted(glob.glob(os.path.join(ROOT, "sigkernel_b200/csrc/build/*.ptxas.log"))):
    s = open(log).read()
    for name, st, ss, sl, regs in re.findall(
            r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", s):
        m = re.search(r"solver_kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb([01])ELi(\d+)E", name)
        if not m:
            continue
        mode, kind, rc, ld, dp2, ex, minb = map(int, m.groups())
        rows.append((MODES[mode], KINDS[kind], rc, ld, dp2, ex, minb, int(regs), int(st), int(ss), int(sl)))
for r in rows:
    line = "%-8s %-6s RC=%d LOGD=%d DP2=%d EX=%d MINB=%-2d regs=%-3d stack=%-4d spill_st=%-4d spill_ld=%d" % r
    if flt in line:
        print(line)
print(len(rows), "instantiations;", sum(1 for r in rows if r[9] > 0), "with spills")
