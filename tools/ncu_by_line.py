#!/usr/bin/env python3
"""Aggregate an ncu SASS-level source page (csv) by CUDA source line using nvdisasm line info.
usage: ncu_by_line.py <sass_page.csv> <nvdisasm -g -c output> <kernel mangled name substring>"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1:4]
# address -> (file, line) from nvdisasm
addr2line = {}
cur = None; infn = False; fileline = ("?", 0)
for l in open(dis, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kname in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        fileline = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", l)
    if m: addr2line[int(m.group(1), 16)] = fileline
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
isrc = hdr.index("Source")
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
base = None
for r in rows[2:]:
    try: a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except Exception: continue
    if base is None: base = a
    fl = addr2line.get(a - base, ("?", 0))
    s = int(r[isamp] or 0); e = int(r[iexec] or 0)
    g = agg[fl]; g[0] += s; g[1] += e; g[2] += 1
    if re.match(r"\s*(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP|MNMX)", r[isrc]): g[3] += e
tot_s = sum(v[0] for v in agg.values()); tot_e = sum(v[1] for v in agg.values()); tot_d = sum(v[3] for v in agg.values())
print(f"total samples {tot_s} inst executed {tot_e} fp64 executed {tot_d}")
for fl, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 60]:
    print(f"{fl[0]}:{fl[1]:<5d} samples {100*v[0]/tot_s:5.2f}%  exec {100*v[1]/tot_e:5.2f}%  fp64 {100*v[3]/max(1,tot_d):5.2f}%  static {v[2]}")
