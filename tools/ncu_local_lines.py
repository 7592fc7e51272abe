#!/usr/bin/env python3
"""Local-memory traffic (spills, demoted arrays) of a kernel by CUDA source line, in doubles per thread per step.
usage: ncu_local_lines.py <sass_page.csv> <nvdisasm -g -c listing> <kernel substring> <members> <steps> [top]"""
import csv, re, sys, collections
sass_csv, dis, kname, members, steps = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
addr2line = {}; infn = False; fl = ("?", 0)
for l in open(dis, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: infn = kname in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: fl = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", l)
    if m: addr2line[int(m.group(1), 16)] = fl
rows = list(csv.reader(open(sass_csv))); hdr = rows[1]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
base = None; agg = collections.Counter(); tot = collections.Counter()
wsteps = int(members) / 32 * int(steps)
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    m = re.search(r"\b(LDL|STL)(\.LU)?(\.(64|128))?", r[isrc])
    if not m: continue
    e = int(r[iex] or 0) * {"64": 2, "128": 4}.get(m.group(4), 1) / 2
    agg[(addr2line.get(a - base, ("?", 0)), m.group(1))] += e; tot[m.group(1)] += e
print("doubles per thread-step:", {k: round(v / wsteps, 1) for k, v in tot.items()})
for (l, op), v in agg.most_common(top): print(f"{l[0]}:{l[1]:<5d} {op}  {v / wsteps:7.2f}")
