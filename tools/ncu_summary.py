#!/usr/bin/env python
"""Markdown summary of an .ncu-rep (headline metrics table + hottest source lines). Source text that ncu could not embed (kernels the
stage specialiser compiled with NVRTC from in-memory headers) is looked up in tuplex_b200/csrc/ by (file, line).
usage: ncu_summary.py <rep> <title> [top-n]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, title = sys.argv[1], sys.argv[2]
topn = sys.argv[3] if len(sys.argv) > 3 else "14"
out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles_tool.py"), rep, topn], capture_output=True, text=True).stdout.splitlines()
print(f"## {title}\n")
print("| metric | value |\n|---|---|")
for kv in out[0].split(" | "):
    k, _, v = kv.partition("=")
    print(f"| {k} | {v} |")
print("\nHottest source lines (share of warp instructions / of stall samples):\n```")
cache = {}
for line in out[1:]:
    if "<Unable to open source file>" in line:
        f, _, rest = line.partition(":")
        f = f.strip()
        try:
            ln = int(rest.split()[0])
            if f not in cache:
                p = os.path.join(ROOT, "tuplex_b200", "csrc", f)
                cache[f] = open(p).read().splitlines() if os.path.exists(p) else None
            txt = cache[f][ln - 1].strip()[:110] if cache[f] else "(generated row function)"
        except (ValueError, IndexError):
            txt = ""
        line = line.replace("<Unable to open source file>", txt)
    print(line)
print("```\n")
