#!/usr/bin/env python
"""For every LDG in the largest-address loop range given, print how many instructions later its destination register is first read.
usage: tools/sass_ldg_distance.py OBJ KERNEL_SUBSTRING LO HI"""
import re, subprocess, sys
obj, key, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        if on: break
        on = key in line
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and lo <= int(m.group(1), 16) <= hi: ins.append((int(m.group(1), 16), m.group(2).strip()))
n = len(ins)
for i, (a, s) in enumerate(ins):
    m = re.search(r"LDG\S*\s+R(\d+),", s)
    if not m: continue
    r = int(m.group(1)); regs = {f"R{r}", f"R{r+1}"} if ".64" in s else {f"R{r}"}
    for d in range(1, n + 1):
        a2, s2 = ins[(i + d) % n]
        ops = s2.split(None, 1)[1] if " " in s2 else ""
        srcs = ops.split(",", 1)[1] if "," in ops else ops
        if s2.startswith(("ST", "@")) or "ST" in s2.split()[0]: srcs = ops
        if any(re.search(rf"\b{x}\b", srcs) for x in regs):
            print(f"{a:x}: {s[:70]:70s} first use +{d:4d} instr at {a2:x}: {s2[:50]}{'  (next iteration)' if i + d >= n else ''}")
            break
