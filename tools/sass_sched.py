#!/usr/bin/env python
"""Decode scheduling control fields (stall count, yield, scoreboard waits) from `cuobjdump -sass` output.
usage: cuobjdump -sass lib.so | python tools/sass_sched.py <function-substring> [first_line last_line]"""
import re
import sys

name = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
inside = False
pend = None
rows = []
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        inside = name in m.group(1)
        continue
    if not inside:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", line)
    if m:
        pend = (m.group(1), m.group(2).strip(), int(m.group(3), 16))
        continue
    m = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", line)
    if m and pend:
        hi64 = int(m.group(1), 16)
        stall = (hi64 >> 41) & 0xf
        yld = (hi64 >> 45) & 1
        wbar = (hi64 >> 46) & 7
        rbar = (hi64 >> 49) & 7
        wmask = (hi64 >> 52) & 0x3f
        rows.append((pend[0], pend[1], stall, yld, wbar, rbar, wmask))
        pend = None
tot = 0
for i, (addr, txt, stall, yld, wbar, rbar, wmask) in enumerate(rows):
    if lo <= i < hi:
        tot += stall
        print(f"{i:5d} {addr} s={stall:2d} y={yld} w={wbar if wbar != 7 else '-'} r={rbar if rbar != 7 else '-'} m={wmask:06b}  {txt[:80]}")
if lo or hi < 10**9:
    print("sum of stall counts in range:", tot, "instructions:", min(hi, len(rows)) - lo)
