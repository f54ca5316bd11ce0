"""One K8 build + a few probes of the bench's join workload (50 M probe rows x 1 M-row table), for ncu captures and kernel timing:
  python tools/join_probe.py            # prints CUDA-event kernel ms per probe
  ncu --set full -k regex:join_ -c 14 ... python tools/join_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuplex_b200 import backend, ir  # noqa: E402
from tuplex_b200.backend import Column  # noqa: E402

n_probe, n_build = int(os.environ.get("PROBE_ROWS", "50000000")), int(os.environ.get("BUILD_ROWS", "1000000"))
rng = np.random.default_rng(42)
bkeys = rng.permutation(n_build).astype(np.int64) * 2 + 1
names = np.frombuffer(b"".join(b"%08d" % i for i in range(n_build)), dtype=np.uint8).copy()
build = [Column(ir.T_I64, bkeys), Column(ir.T_I64, rng.integers(0, 1 << 40, n_build)),
         Column(ir.T_STR, names, (np.arange(n_build + 1, dtype=np.uint64) * 8).astype(np.uint32))]
pkeys = bkeys[rng.integers(0, n_build, n_probe)]
miss = rng.random(n_probe) < 0.1
pkeys[miss] += 1
probe = [Column(ir.T_I64, pkeys), Column(ir.T_I64, rng.integers(0, 1 << 40, n_probe))]
backend.init([0])
bb = backend.Block.upload(0, build, n_build)
pb = backend.Block.upload(0, probe, n_probe)
jn = backend.Join(bb, [c.type for c in build], 0)
print("build:", jn.info)
for k in range(int(os.environ.get("PROBE_REPEAT", "3"))):
    t0 = time.perf_counter()
    res = jn.probe(pb, [c.type for c in probe], 0)
    info = res.info
    print("probe %d: kernel %.3f ms, wall %.3f ms, %d out rows, %d launches -> %.2f G probe rows/s"
          % (k, info.kernel_ms, (time.perf_counter() - t0) * 1e3, info.n_out_rows, info.kernel_launches, n_probe / info.kernel_ms / 1e6))
    res.free()
