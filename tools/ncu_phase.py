#!/usr/bin/env python3
"""Split an integrator kernel's ncu SASS page into per-RHS code (executed ~4x per step), once-per-step code (callbacks,
control laws) and rare code, with the stall mix of each; then the top once-per-step source lines.
usage: ncu_phase.py <sass_page.csv> <nvdisasm -g -c listing> <kernel substring> <members> <steps> [top]"""
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
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
cols = [c for c in hdr if c.startswith("stall_") and "Not" not in c]
W = int(members) / 32 * int(steps)
cls = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
once = collections.Counter(); base = None
for r in rows[2:]:
    try: a = int(r[ia], 16); e = int(r[iex] or 0); s = int(r[isamp] or 0)
    except Exception: continue
    if base is None: base = a
    ratio = e / W
    k = "per RHS (~4/step)" if ratio > 3.5 else ("once per step" if 0.9 < ratio <= 1.2 else ("2-3.5/step" if ratio > 1.2 else "rare (<0.9/step)"))
    c = cls[k]; c[0] += s; c[1] += e; c[2] += 1
    for col in cols: c[3][col] += int(r[hdr.index(col)] or 0)
    if k == "once per step": once[addr2line.get(a - base, ("?", 0))] += s
tot = sum(c[0] for c in cls.values())
for k, c in cls.items():
    print(f"{k:18s} samples {100 * c[0] / tot:5.1f}%  warp-instr/step {c[1] / W:7.0f}  static {c[2]:6d}  " +
          ", ".join(f"{a[6:]} {100 * b / max(c[0], 1):.0f}%" for a, b in c[3].most_common(5)))
print("once-per-step lines (share of all samples):")
for fl, v in once.most_common(top): print(f"  {fl[0]}:{fl[1]:<5d} {100 * v / tot:5.2f}%")
