import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from tuplex_b200 import backend, frontend, ir
from oracle import pyoracle
import scan_udfs as U
backend.init([0])
n = 50001
sc = frontend.StageCompiler(U.TYPES, U.NAMES)
U._h_zillow(sc)
prog = sc.finish_memory(prefilter=False)   # the head alone as a row stage: outputs = all columns + bedrooms
pre = frontend.StageCompiler(U.TYPES, U.NAMES); U._h_zillow(pre); U.heavy_tail(pre, 100100); full = pre.finish_memory()
cols = U.make_columns(n, n)
ora = pyoracle.run_program(full, cols, n, 4)
vals = [c.to_values() for c in cols]
for env in ({}, {"TPLX_MASK_STAGE": "0"}, {"TPLX_NO_SCAN": "1"}):
    for k in ("TPLX_MASK_STAGE", "TPLX_NO_SCAN"): os.environ.pop(k, None)
    os.environ.update(env)
    res = backend.Stage(full).run_host(0, cols, n, 4)
    print(env, "rows", int(res.info.n_out_rows), "oracle", ora.n_out, "exc", int(res.info.n_exceptions), len(ora.exceptions))
    # which input rows are missing? the last user column of the oracle has no row index; recompute kept rows in python
    got_t = res.column(1).to_values()  # title column passes through
    exp_t = ora.values(1)
    if got_t != exp_t:
        # align greedily
        i = j = 0; miss = []
        while j < len(exp_t) and len(miss) < 10:
            if i < len(got_t) and got_t[i] == exp_t[j] and res.column(0).to_values()[i] == ora.values(0)[j]: i += 1; j += 1
            else: miss.append((j, ora.values(0)[j], exp_t[j])); j += 1
        print("  missing examples (oracle out idx, facts, title):", miss)
