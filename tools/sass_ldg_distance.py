#!/usr/bin/env python
"""How many instructions before its first use is each global load of a kernel issued?

ptxas places a plain (and, less freely, an ordered) global load wherever it likes between its address and its first use; in the sweep
kernels that distance decides whether the load's latency is covered (builds with < 30 instructions between the
per-step loads and their shared-memory deposits ran 3-7x slower with bit-identical results).

usage: tools/sass_ldg_distance.py OBJ KERNEL_SUBSTRING [LO HI]     (addresses in hex; default: the whole kernel, no wrap-around)
"""
import re
import subprocess
import sys


def dump_sass(obj):
    return subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout


def kernel_sass(obj, key, txt=None):
    """[(address, text)] of the first kernel whose mangled name contains `key` (txt: a dump_sass() result to reuse)."""
    if txt is None:
        txt = dump_sass(obj)
    ins, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            if on:
                break
            on = key in line
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def ldg_distances(ins, wrap=False):
    """For every LDG: (address, text, distance in instructions to the first reader of its destination, that reader's text)."""
    out, n = [], len(ins)
    for i, (a, s) in enumerate(ins):
        m = re.search(r"LDG\S*\s+R(\d+),", s)
        if not m:
            continue
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        r = int(m.group(1))
        regs = {f"R{r}", f"R{r + 1}"} if ".64" in op else {f"R{r}"}
        for d in range(1, (n if wrap else n - i)):
            a2, s2 = ins[(i + d) % n]
            body = re.sub(r"^@!?U?P\w+\s+", "", s2)
            ops = body.split(None, 1)[1] if " " in body else ""
            is_store = body.split()[0].startswith(("ST", "RED", "ATOM"))
            srcs = ops if is_store else (ops.split(",", 1)[1] if "," in ops else "")
            dst = "" if is_store else ops.split(",", 1)[0]
            if any(re.search(rf"\b{x}\b", srcs) for x in regs):
                out.append((a, s, d, s2))
                break
            if any(re.search(rf"\b{x}\b", dst) for x in regs):      # overwritten before any read: a dead or predicated-off value
                break
    return out


if __name__ == "__main__":
    obj, key = sys.argv[1], sys.argv[2]
    ins = kernel_sass(obj, key)
    wrap = False
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
        ins = [x for x in ins if lo <= x[0] <= hi]
        wrap = True
    for a, s, d, s2 in ldg_distances(ins, wrap):
        print(f"{a:x}: {s[:70]:70s} first use +{d:4d} instr: {s2[:50]}")
