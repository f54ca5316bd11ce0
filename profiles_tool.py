#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics and the hottest source lines (needs -lineinfo)."""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    print(" | ".join(f"{w.split('.')[0].replace('launch__','').replace('smsp__','').replace('sm__','')}={r[hdr.index(w)]}" for w in want if w in hdr))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = defaultdict(lambda: [0, 0]); text = {}; cur = None; f = None; kern = -1; per_kernel = defaultdict(lambda: defaultdict(lambda: [0, 0]))
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Kernel Name": kern += 1; continue
    if r[0] == "Line No": continue
    if r[0] != "":
        try: cur = (f, int(r[0])); text[cur] = r[1][:100]
        except ValueError: pass
        continue
    if len(r) > 7 and r[2].startswith("0x"):
        try: n = int(r[7]); s = int(r[6])
        except ValueError: continue
        agg[cur][0] += n; agg[cur][1] += s
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions (all profiled launches): {tot}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{k[0]:14s}:{k[1]:4d} inst {v[0]/tot*100:5.1f}%  samples {v[1]/tots*100:5.1f}% | {text.get(k, '')}")
