#!/usr/bin/env python
"""Print the instructions of one kernel between two addresses, optionally without the FP64 arithmetic.
usage: tools/sass_dump.py OBJ KERNEL_SUBSTRING LO HI [nofp64]"""
import re, subprocess, sys
obj, key, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
nof = len(sys.argv) > 5
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on = False
for line in txt.splitlines():
    if "Function :" in line:
        if on: break
        on = key in line
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and lo <= int(m.group(1), 16) <= hi and not (nof and re.search(r"\b(DFMA|DMUL|DADD)\b", m.group(2))):
        print(m.group(1), m.group(2).strip())
