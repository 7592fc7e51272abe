#!/usr/bin/env python3
"""Source lines ranked by one stall reason, from an ncu SASS source page (csv) + nvdisasm -g -c listing.
usage: ncu_stall_lines.py <sass_page.csv> <nvdisasm listing> <kernel substring> <stall column, e.g. stall_long_sb> [top]"""
import csv, re, sys, collections
sass_csv, dis, kname, col = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
addr2line = {}; infn = False; fileline = ("?", 0)
for l in open(dis, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: infn = kname in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: fileline = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", l)
    if m: addr2line[int(m.group(1), 16)] = fileline
rows = list(csv.reader(open(sass_csv))); hdr = rows[1]
ia, isrc, ic, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index(col), hdr.index("# Samples")
agg = collections.Counter(); ops = collections.defaultdict(collections.Counter); base = None; tot = 0; alls = 0
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    v = int(r[ic] or 0); tot += v; alls += int(r[isamp] or 0)
    fl = addr2line.get(a - base, ("?", 0)); agg[fl] += v
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc]); ops[fl][m.group(2) if m else "?"] += v
print(f"{col}: {tot} of {alls} samples ({100*tot/alls:.1f}%)")
for fl, v in agg.most_common(top):
    print(f"{fl[0]}:{fl[1]:<5d} {100*v/tot:5.2f}%   " + ", ".join(f"{o} {c}" for o, c in ops[fl].most_common(3)))
