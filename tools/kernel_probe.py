"""Kernel-time probe: Zillow stage over one device-resident block; prints CUDA-event kernel ms (prefilter+dense)."""
import sys, os
os.environ.setdefault("TPLX_JIT_SYNC", "1")  # steady state: wait for the stage specialiser at the first block
sys.path.insert(0, '/root/repo')
import numpy as np
from tuplex_b200 import backend, workloads as W
backend.init([0])
src, n0 = W.load_zillow_fixture()
n = int(os.environ.get("PROBE_CYCLES", "400")) * n0
cols = W.replicate(src, n0, n)
st = backend.Stage(W.zillow_program())
blk = backend.Block.upload(0, cols, n)
ms = []
for it in range(6):
    if it == 5: os.environ["TPLX_TRACE"] = "1"
    r = st.run(blk); inf = r.info; ms.append(inf.kernel_ms); r.free()
print(os.environ.get("TPLX_GPU_LIB", "default").split("/")[-1], "rows", n, "kernel ms", ["%.3f" % m for m in ms[2:]], "-> %.2f G rows/s" % (n / (min(ms[2:]) * 1e-3) / 1e9))
