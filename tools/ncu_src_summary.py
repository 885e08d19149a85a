#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): stall reasons and the hottest instructions.
usage: ncu -i rep --page source --csv | python tools/ncu_src_summary.py [top_n]"""
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
S = col["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in data)
print(f"instructions: {len(data)}  samples: {tot}")
for h in sorted(stalls, key=lambda h: -sum(int(r[col[h]]) for r in data)):
    v = sum(int(r[col[h]]) for r in data)
    if v:
        print(f"  {h:24s} {v:8d} {100.0 * v / tot:5.1f}%")
op = {}
for r in data:
    name = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if name.startswith("@"):
        name = r[col["Source"]].split()[1]
    name = name.split(".")[0]
    e = op.setdefault(name, [0, 0, 0])
    e[0] += 1; e[1] += int(r[S]); e[2] += int(r[col["Instructions Executed"]])
print("by opcode (static count, samples, warp-instr executed):")
for k, e in sorted(op.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {k:10s} n={e[0]:5d} samples={e[1]:8d} ({100.0 * e[1] / tot:5.1f}%) exec={e[2]}")
print(f"top {top_n} instructions by samples:")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:top_n]
for i in sorted(idx):
    r = data[i]
    why = sorted(((int(r[col[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"  [{i:5d}] {int(r[S]):6d}  {r[col['Source']].strip()[:70]:70s} {why[0][1]}={why[0][0]} {why[1][1]}={why[1][0]}")
