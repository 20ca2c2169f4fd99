#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line using nvdisasm -g line info.
usage: ncu_by_line.py <source.csv> <nvdisasm -g output> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
addr2line = {}
cur = ("?", 0)
nsec = 0
for l in open(sass, errors="replace"):
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        nsec += 1
        if nsec > 1 and addr2line:
            break  # only the first function of the listing
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
base = None
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    key = addr2line.get(a - base, (("?", 0), ""))[0]
    for name in ["# Samples", "Instructions Executed"] + stalls:
        try:
            v = float(r[col[name]])
        except ValueError:
            v = 0
        agg[key][name] += v
        tot[name] += v
print("total samples", tot["# Samples"], "instructions", tot["Instructions Executed"])
print("stall mix:", {s: round(100 * tot[s] / max(tot["# Samples"], 1), 1) for s in stalls if tot[s] > 0.01 * tot["# Samples"]})
for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    mix = sorted(((d[s], s) for s in stalls), reverse=True)[:3]
    print(f"{key[0]}:{key[1]:4d}  samples {100 * d['# Samples'] / tot['# Samples']:5.1f}%  inst {100 * d['Instructions Executed'] / tot['Instructions Executed']:5.1f}%  "
          + " ".join(f"{s[6:]}={100 * v / max(d['# Samples'], 1):.0f}%" for v, s in mix))
