"""Diagnostic: per-tile phase clocks of the specialised wide-tile kernel (TPLX_JIT_TIMES=<file>), C1 over PROBE_ROWS rows.
Prints the median / p90 duration of every phase of a tile in microseconds (SM clock 1.965 GHz)."""
import os, sys
os.environ.setdefault("TPLX_JIT_SYNC", "1")
os.environ["TPLX_JIT_TIMES"] = os.environ.get("TPLX_JIT_TIMES", "/tmp/tile_times.bin")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from tuplex_b200 import backend, workloads as W
from tuplex_b200.backend import Column
backend.init([0])
n = int(os.environ.get("PROBE_ROWS", "100000000"))
x = np.arange(1, n + 1, dtype=np.int64)
st = backend.Stage(W.c1_program())
blk = backend.Block.upload(0, [Column(0, x)], n)
for it in range(4):
    r = st.run(blk); ms = r.info.kernel_ms; r.free()
t = np.fromfile(os.environ["TPLX_JIT_TIMES"], dtype=np.uint64).reshape(-1, 8)
t = t[t[:, 0] != 0]
ghz = 1.965
names = ["load+eval", "counts+scan+publish", "look-back", "stores", "next ticket (atomic)"]
print("kernel ms %.3f, tiles %d, CTAs %d" % (ms, len(t), len(np.unique(t[:, 7]))))
for i, nm in enumerate(names):
    d = (t[:, i + 1].astype(np.int64) - t[:, i].astype(np.int64)) / ghz / 1e3
    print("%-24s median %6.2f us  p90 %6.2f us  mean %6.2f us" % (nm, np.median(d), np.percentile(d, 90), d.mean()))
tot = (t[:, 5].astype(np.int64) - t[:, 0].astype(np.int64)) / ghz / 1e3
print("%-24s median %6.2f us  p90 %6.2f us  mean %6.2f us" % ("whole tile", np.median(tot), np.percentile(tot, 90), tot.mean()))
# gap between a CTA's consecutive tiles (barrier at the loop top + whatever else)
order = np.lexsort((t[:, 0], t[:, 7]))
s = t[order]
same = s[1:, 7] == s[:-1, 7]
gap = (s[1:, 0].astype(np.int64) - s[:-1, 5].astype(np.int64))[same] / ghz / 1e3
print("%-24s median %6.2f us  p90 %6.2f us" % ("gap to the next tile", np.median(gap), np.percentile(gap, 90)))
