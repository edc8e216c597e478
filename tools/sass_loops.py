#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in an object file with their size and opcode mix.
usage: tools/sass_loops.py OBJ KERNEL_SUBSTRING [MIN_INSTR]"""
import collections, re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = key in line
        if on and ins: break
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print(len(ins), "instructions")
for a, s in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\w+,\s*)?0x([0-9a-f]+)", s)
    if m and int(m.group(1), 16) < a:
        lo = int(m.group(1), 16)
        body = [x for x in ins if lo <= x[0] <= a]
        if len(body) < minlen: continue
        h = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
        fp64 = sum(h[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print(f"loop 0x{lo:x}..0x{a:x}: {len(body)} instr, fp64 {fp64}:", ", ".join(f"{k} {v}" for k, v in h.most_common(24)))
