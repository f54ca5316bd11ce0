"""One CSV block (250 cycles of the Zillow fixture file, 1.675 GB) parsed twice + the Z1 stage: the command profiled by
ncu for profiles/r01_csv_*.  usage: python tools/csv_probe.py [cycles]"""
import gzip
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from tuplex_b200 import backend, ir, workloads as W  # noqa: E402

cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 250
raw = gzip.open(os.path.join(ROOT, "tests", "golden", "zillow_noexc.csv.gz"), "rb").read()
header, body = raw.split(b"\n", 1)
text = np.frombuffer(header + b"\n" + body * cycles, dtype=np.uint8)
backend.init([0])
S, F, X = ir.T_STR, ir.T_F64, backend.CSV_SKIP
types = [S, S, S, S, F, S, S, X, S, X]
buf = backend.CsvBuffer(0, text)
st = backend.Stage(W.zillow_program())
for it in range(2):
    t0 = time.perf_counter()
    p = buf.parse(types, header=True)
    r = st.run(p.block, 0)
    print("rows", int(p.info.n_rows), "bad", int(p.info.n_bad), "parse_ms", round(p.info.parse_ms, 3), "stage_ms", round(r.info.kernel_ms, 3),
          "out", int(r.info.n_out_rows), "wall_ms", round((time.perf_counter() - t0) * 1e3, 2))
    r.free()
    p.free()

# host -> device bandwidth of the text itself (page-locked source), for the e2e discussion in profiles/
import torch  # noqa: E402
t = torch.empty(text.size, dtype=torch.uint8, pin_memory=True)
t.numpy()[...] = text
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    bs = [backend.CsvBuffer(0, t.numpy()) for _ in range(3)]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("upload x3: %.1f ms, %.1f GB/s" % (dt * 1e3, 3 * text.size / dt / 1e9))
    for b in bs:
        b.free()
