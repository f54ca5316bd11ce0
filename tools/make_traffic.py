#!/usr/bin/env python
"""profiles/traffic.json from the `ncu --set full` captures of tools/r02_capture.sh: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum)
per input row of the dominant kernels, keyed by the hash of the stage-kernel sources they were captured from (bench.kernel_source_hash):
bench.py reports `roofline.traffic` from it only while that hash still matches.
usage: make_traffic.py <dir with r02_scan / r02_dense / r02_vec / r02_q6 .ncu-rep>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def dram_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        tot += float(vals[i].replace(",", "")) * scale
    return tot, vals[hdr.index("Kernel Name")]


d = sys.argv[1]
zill_rows, c1_rows, q6_rows = 32661 * int(os.environ.get("PROBE_CYCLES", "400")), int(os.environ.get("PROBE_ROWS", "50000000")), 100_000_000
out = {"kernel_source_hash": bench.kernel_source_hash(), "dram_bytes_per_row": {}, "captures": {}}
scan, dense = os.path.join(d, "r02_scan.ncu-rep"), os.path.join(d, "r02_dense.ncu-rep")
if os.path.exists(scan) and os.path.exists(dense):
    a, ka = dram_bytes(scan)
    b, kb = dram_bytes(dense)
    out["dram_bytes_per_row"]["zillow_z1"] = (a + b) / zill_rows
    out["captures"]["zillow_z1"] = {"rows": zill_rows, "kernels": {ka: a, kb: b}, "note": "prefilter + dense launch of one block; mask_count/scan/expand (< 1 %) not included"}
vec = os.path.join(d, "r02_vec.ncu-rep")
if os.path.exists(vec):
    a, ka = dram_bytes(vec)
    out["dram_bytes_per_row"]["c1_map_filter"] = a / c1_rows
    out["captures"]["c1_map_filter"] = {"rows": c1_rows, "kernels": {ka: a}}
q6 = os.path.join(d, "r02_q6.ncu-rep")
if os.path.exists(q6):
    a, ka = dram_bytes(q6)
    out["dram_bytes_per_row"]["tpch_q6"] = a / q6_rows
    out["captures"]["tpch_q6"] = {"rows": q6_rows, "kernels": {ka: a}}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out["dram_bytes_per_row"]))
