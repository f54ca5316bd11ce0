"""Kernel-time probe: C1 (map x*x, filter even) over one device-resident block of 100M i64 rows through K1v."""
import sys, os
os.environ.setdefault("TPLX_JIT_SYNC", "1")  # steady state: wait for the stage specialiser at the first block
sys.path.insert(0, '/root/repo')
import numpy as np
from tuplex_b200 import backend, workloads as W
from tuplex_b200.backend import Column
backend.init([0])
n = int(os.environ.get("PROBE_ROWS", "100000000"))
x = np.arange(1, n + 1, dtype=np.int64)
st = backend.Stage(W.c1_program())
blk = backend.Block.upload(0, [Column(0, x)], n)
ms = []
for it in range(6):
    r = st.run(blk); inf = r.info; ms.append(inf.kernel_ms); no = int(inf.n_out_rows); r.free()
print("c1 rows", n, "out", no, "kernel ms", ["%.3f" % m for m in ms[2:]], "-> %.1f G rows/s, %.0f GB/s of 12 B/row" % (n / (min(ms[2:]) * 1e-3) / 1e9, n * 12 / (min(ms[2:]) * 1e-3) / 1e9))
