#!/usr/bin/env python3
"""Once-per-step code of an integrator kernel by CUDA source line: warp instructions per step, samples, top stall.
usage: ncu_once_lines.py <sass_page.csv> <nvdisasm -g -c listing> <kernel substring> <members> <steps> [top]"""
import csv, re, sys, collections
sass_csv, dis, kname, members, steps = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
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
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
cols = [c for c in hdr if c.startswith("stall_") and "Not" not in c]
W = int(members) / 32 * int(steps)
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()]); base = None; tot = 0
for r in rows[2:]:
    try: a = int(r[ia], 16); e = int(r[iex] or 0); s = int(r[isamp] or 0)
    except Exception: continue
    if base is None: base = a
    tot += s
    ratio = e / W
    if not (0.9 < ratio <= 1.2): continue
    g = agg[addr2line.get(a - base, ("?", 0))]; g[0] += s; g[1] += e; g[2] += 1
    for col in cols: g[3][col] += int(r[hdr.index(col)] or 0)
print("line, samples % of kernel, warp-instr per step, top stalls")
for fl, g in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{fl[0]}:{fl[1]:<5d} samples {100 * g[0] / tot:5.2f}%  instr/step {g[1] / W:6.1f}  " + ", ".join(f"{a[6:]} {100 * b / max(g[0], 1):.0f}%" for a, b in g[3].most_common(3)))
