"""Instruction mix of an address range of a kernel's SASS (development aid).
usage: python tools/sass_mix.py <lib.so> <kernel-substring> [lo_hex hi_hex]"""
import collections
import re
import subprocess
import sys


def main():
    lib, sub = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    on = False
    mix = collections.Counter()
    n = 0
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
        ins = m.group(2).split()
        op = ins[1] if ins[0].startswith("@") else ins[0]
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD") and "." in op else "")
        mix[op] += 1
        n += 1
    dp = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD"))
    print(f"{n} instructions, {dp} DP ({100.0*dp/max(n,1):.1f}%)")
    for k, v in mix.most_common(40):
        print(f"  {v:5d} {k}")


if __name__ == "__main__":
    main()
