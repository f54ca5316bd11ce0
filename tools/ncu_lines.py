#!/usr/bin/env python
"""Per-kernel hottest source lines of an .ncu-rep (needs -lineinfo): python tools/ncu_lines.py rep [kernel_index] [topn]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
seen = set(); kern = -1; f = None; cur = None; text = {}
agg = defaultdict(lambda: [0, 0])
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0] == "Kernel Name": continue
    if r[0] == "File Path":
        f = r[1].split("/")[-1]
        if kern < 0 or r[1] in seen: kern += 1; seen = set()
        seen.add(r[1]); continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] != "":
        try: cur = (f, int(r[0])); text[cur] = r[1][:110]
        except ValueError: pass
        continue
    if kern == which and len(r) > 7 and r[2].startswith("0x"):
        try: n = int(r[7]); s = int(r[6])
        except ValueError: continue
        agg[cur][0] += n; agg[cur][1] += s
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"kernel #{which}: total warp instructions {tot}")
byfile = defaultdict(int)
for k, v in agg.items(): byfile[k[0]] += v[0]
print("by file:", {k: f"{v/tot*100:.1f}%" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])})
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{k[0]:14s}:{k[1]:4d} inst {v[0]/tot*100:5.1f}%  samples {v[1]/tots*100:5.1f}% | {text.get(k, '')}")
