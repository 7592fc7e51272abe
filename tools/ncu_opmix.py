#!/usr/bin/env python3
"""Dynamic opcode mix and stall attribution from an ncu SASS source page (csv).
usage: ncu_opmix.py <sass_page.csv> [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iexec, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
mix = collections.defaultdict(lambda: [0, 0, 0]); stall_by_op = collections.defaultdict(lambda: collections.Counter())
tot_e = tot_s = 0
stalls = collections.Counter()
for r in rows[2:]:
    if len(r) <= isamp: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", r[isrc])
    if not m: continue
    op = m.group(2)
    full = op + (m.group(3) or "")
    if op == "IMAD" and ".MOV" in full: op = "IMAD.MOV"
    e = int(r[iexec] or 0); s = int(r[isamp] or 0)
    g = mix[op]; g[0] += e; g[1] += s; g[2] += 1
    tot_e += e; tot_s += s
    for i, h in stall_cols:
        v = int(r[i] or 0); stalls[h] += v; stall_by_op[op][h] += v
print(f"warp instructions executed {tot_e}  samples {tot_s}")
print("stalls:", ", ".join(f"{h[6:]} {100*v/tot_s:.1f}%" for h, v in stalls.most_common(10)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for op, (e, s, n) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:top]:
    ms = ", ".join(f"{h[6:]} {100*v/max(1,s):.0f}%" for h, v in stall_by_op[op].most_common(3))
    print(f"{op:10s} exec {100*e/tot_e:5.2f}%  samples {100*s/tot_s:5.2f}%  static {n:5d}   [{ms}]")
