import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tuplex_b200 import backend, workloads as W
from tuplex_b200.backend import Column
backend.init([0])
src, n0 = W.load_zillow_fixture()
n = 500 * n0
cols = W.replicate(src, n0, n)
def pin(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); v = t.numpy(); v[...] = a; return v, t
keep = []
pc = []
for c in cols:
    d, k = pin(c.data); keep.append(k)
    o = None
    if c.offsets is not None:
        o, k2 = pin(c.offsets); keep.append(k2)
    pc.append(Column(c.type, d, o))
st = backend.Stage(W.zillow_program())
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = st.run_host(0, pc, n)
    inf = r.info
    t1 = time.perf_counter()
    outs = r.columns(); ex = r.exceptions()
    t2 = time.perf_counter()
    print(f"run_host {1e3*(t1-t0):.2f} ms (events: total {inf.total_ms:.2f}, kernels {inf.kernel_ms:.2f}), fetch {1e3*(t2-t1):.2f} ms, h2d {inf.h2d_bytes/1e6:.0f} MB zc {inf.zero_copy_cols}")
    r.free()
# resident for comparison
blk = backend.Block.upload(0, pc, n)
for it in range(3):
    r = st.run(blk); inf = r.info; print(f"resident: total {inf.total_ms:.2f} kernels {inf.kernel_ms:.2f}"); r.free()
