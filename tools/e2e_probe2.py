"""Timeline probe of the pipelined end-to-end path (two or three blocks in flight)."""
import sys, time, threading
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from tuplex_b200 import backend, workloads as W
from tuplex_b200.backend import Column
backend.init([0])
src, n0 = W.load_zillow_fixture()
n = 500 * n0
cols = W.replicate(src, n0, n)
def pin(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); v = t.numpy(); v[...] = a; return v, t
keep = []; pc = []
for c in cols:
    d, k = pin(c.data); keep.append(k); o = None
    if c.offsets is not None:
        o, k2 = pin(c.offsets); keep.append(k2)
    pc.append(Column(c.type, d, o))
st = backend.Stage(W.zillow_program())
T0 = [0.0]
def one(i):
    t0 = time.perf_counter()
    r = st.run_host(0, pc, n, 0)
    inf = r.info
    t1 = time.perf_counter()
    outs = r.columns(); r.exceptions()
    t2 = time.perf_counter()
    r.free()
    return (i, t0 - T0[0], t1 - T0[0], t2 - T0[0], inf.total_ms, inf.kernel_ms)
for workers in (1, 2, 3):
    pool = ThreadPoolExecutor(max_workers=workers)
    list(pool.map(one, range(2)))  # warm
    torch.cuda.synchronize(); T0[0] = time.perf_counter()
    res = list(pool.map(one, range(8)))
    torch.cuda.synchronize(); dt = time.perf_counter() - T0[0]
    print(f"workers={workers}: {8*n/dt/1e6:.0f} M rows/s, {dt/8*1e3:.1f} ms/block")
    for r in res: print("   blk %d start %.1f run_done %.1f fetched %.1f | events total %.1f kernels %.1f" % (r[0], r[1]*1e3, r[2]*1e3, r[3]*1e3, r[4], r[5]))
