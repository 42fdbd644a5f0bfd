"""Estimate fp64-pipe occupancy of a SASS address range under the measured operand-read model
(DADD/DMUL/DFMA occupy the pipe max(2, #distinct 64-bit register sources) cycles).
usage: python tools/sass_dp_cost.py <lib.so> <kernel-substring> [lo_hex hi_hex]"""
import re
import subprocess
import sys


def main():
    lib, sub = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    on = False
    n_instr = n_dp = 0
    cyc_min = cyc_model = cyc_reuse = 0
    hist = {}
    for line in txt.splitlines():
        if "Function :" in line:
            on = sub in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if addr < lo or addr > hi:
            continue
        n_instr += 1
        ins = m.group(2)
        toks = ins.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        base = op.split(".")[0]
        if base not in ("DFMA", "DADD", "DMUL", "DSETP"):
            continue
        n_dp += 1
        ops = ins.split(op, 1)[1]
        parts = [x.strip() for x in ops.split(",")]
        srcs = parts[1:] if base != "DSETP" else parts[2:]
        regs = []
        reuse = 0
        for s_ in srcs:
            mm = re.match(r"[-|]*\s*(R\d+)(\.reuse)?", s_)
            if mm and mm.group(1) != "RZ":
                if mm.group(1) not in regs:
                    regs.append(mm.group(1))
                    if mm.group(2):
                        reuse += 1
        k = len(regs)
        hist[(base, k)] = hist.get((base, k), 0) + 1
        cyc_min += 2
        cyc_model += max(2, k)
        cyc_reuse += max(2, k - reuse)
    print(f"{n_instr} instructions, {n_dp} DP; pipe cycles: 2/instr {cyc_min}, distinct-operand model {cyc_model}, "
          f"if every .reuse source were free {cyc_reuse}")
    for k in sorted(hist):
        print("  ", k, hist[k])


if __name__ == "__main__":
    main()
