#!/usr/bin/env python3
"""Static SASS instruction count of one kernel by section of rhs() (fb_physics.cuh), from `nvdisasm -gi -c` (inline chains).
usage: sass_sections.py <nvdisasm -gi output> <kernel mangled-name substring> [by-line]
Every instruction is attributed to the frame of its inline chain that lies inside the body of rhs() (or, failing that, to
its outermost frame's file); sections are the `// ---- name` banners of rhs()."""
import re, sys, collections, os
dis, kname = sys.argv[1:3]
byline = len(sys.argv) > 3
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flight.jl_b200", "csrc", "fb_physics.cuh")
src = open(root).read().splitlines()
# rhs() body range and section banners
start = next(i for i, l in enumerate(src) if l.startswith("FB_DEV void rhs(")) + 1
end = next(i for i, l in enumerate(src) if i > start and l.startswith("}")) + 1
banners = [(i + 1, re.sub(r"[-/ ]+", " ", l).strip()) for i, l in enumerate(src) if start <= i + 1 <= end and re.match(r"\s*// -{20,}", l)]
def section(line):
    name = "rhs prologue"
    for ln, nm in banners:
        if line >= ln: name = nm
    return name
infn = False; frames = []; pending = []
agg = collections.defaultdict(lambda: [0, 0]); lines = collections.defaultdict(lambda: [0, 0])
for l in open(dis, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: infn = kname in m.group(1); continue
    if not infn: continue
    if re.match(r"\$\S+\$\S+:", l): infn = False; continue      # out-of-line device functions behind the kernel body
    m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3): pending.append((os.path.basename(m.group(4)), int(m.group(5))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        if pending: frames = pending; pending = []
        ins = m.group(2)
        key = None
        for f, ln in frames:
            if f == "fb_physics.cuh" and start <= ln <= end: key = ("rhs: " + section(ln), ln); break
        if key is None:
            f, ln = frames[-1] if frames else ("?", 0)
            key = (f, ln)
            src_k = open(os.path.join(os.path.dirname(root), f)).read().splitlines() if f.endswith((".cuh", ".cu")) and os.path.exists(os.path.join(os.path.dirname(root), f)) else None
        fp64 = bool(re.match(r"(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP)", ins))
        a = agg[key[0]]; a[0] += 1; a[1] += fp64
        b = lines[(key[0], key[1])]; b[0] += 1; b[1] += fp64
tot = sum(v[0] for v in agg.values()); totd = sum(v[1] for v in agg.values())
print(f"static instructions {tot}, FP64 {totd}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:50s} {v[0]:6d} {100*v[0]/tot:5.1f}%   fp64 {v[1]:6d} {100*v[1]/max(1,totd):5.1f}%")
if byline:
    print()
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:80]:
        print(f"{k[0]:40s}:{k[1]:<5d} {v[0]:6d}  fp64 {v[1]:6d}   {src[k[1]-1].strip()[:110] if k[0].startswith('rhs') else (open(os.path.join(os.path.dirname(root), k[0])).read().splitlines()[k[1]-1].strip()[:110] if os.path.exists(os.path.join(os.path.dirname(root), k[0])) and k[1] > 0 else '')}")
