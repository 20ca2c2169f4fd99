#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line using `nvdisasm -g` line info.

usage: ncu_by_line.py <source.csv> <nvdisasm -g listing> <function substring> [section index] [top]
  source.csv : ncu -i rep.ncu-rep --page source --csv   (one "Address" table per captured launch)
  listing    : nvdisasm -g <cubin extracted with cuobjdump -xelf all lib.so>
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, func = sys.argv[1], sys.argv[2], sys.argv[3]
section = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
addr2line = {}
cur = ("?", 0)
inside = False
for l in open(sass, errors="replace"):
    if l.startswith(".text."):
        inside = func in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[section]
end = heads[section + 1] - 1 if section + 1 < len(heads) else len(rows)
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
pipe = defaultdict(float)
base = None
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:end]:
    if len(r) < len(hdr):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    key, text = addr2line.get(a - base, (("?", 0), ""))
    for name in ["# Samples", "Instructions Executed"] + stalls:
        try:
            v = float(r[col[name]])
        except ValueError:
            v = 0
        agg[key][name] += v
        tot[name] += v
    op = r[col["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    if op:
        try:
            pipe[op[0].split(".")[0]] += float(r[col["Instructions Executed"]])
        except ValueError:
            pass
print("total samples", tot["# Samples"], "instructions", tot["Instructions Executed"])
print("stall mix:", {s: round(100 * tot[s] / max(tot["# Samples"], 1), 1) for s in stalls if tot[s] > 0.01 * tot["# Samples"]})
print("opcode mix:", {k: round(100 * v / tot["Instructions Executed"], 1) for k, v in sorted(pipe.items(), key=lambda kv: -kv[1])[:24]})
for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    mix = sorted(((d[s], s) for s in stalls), reverse=True)[:3]
    print(f"{key[0]}:{key[1]:4d}  samples {100 * d['# Samples'] / tot['# Samples']:5.1f}%  inst {100 * d['Instructions Executed'] / tot['Instructions Executed']:5.1f}%  "
          + " ".join(f"{s[6:]}={100 * v / max(d['# Samples'], 1):.0f}%" for v, s in mix))
