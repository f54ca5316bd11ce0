import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from tuplex_b200 import backend, frontend
from oracle import pyoracle
from fuzz_udfs import COLS, TYPES, apply_ops, make_columns
backend.init([0])
seed = 23; n = 20000 + 37 * seed
cols, _ = make_columns(n, seed)
ops = [('add_with_column', 'c0', "lambda x: (int(x['t'][-1:-1]) ^ (-(x['a'] - x['a'])))"), ('add_with_column', 'c1', "lambda x: (float(x['b']) * float(len(x['s'])))"), ('add_filter', 'lambda x: (6 != -4)'), ('add_with_column', 'c2', "lambda x: ((min(x['b'], len(x['s'])) - x['s'].find('HO')) & (' 7'.find('a') if (0.1 == 0.1) else int(x['s'])))"), ('add_filter', "lambda x: (x['b'] > len(x['s']))")]
sc = frontend.StageCompiler(TYPES, COLS); apply_ops(sc, ops); prog = sc.finish_memory(prefilter=True)
print(prog.prefilter.dump())
ora = pyoracle.run_program(prog, cols, n, first_row_no=seed)
print("oracle exc", len(ora.exceptions), ora.exceptions[:3], ora.exceptions[-2:])
for env in ({}, {"TPLX_NO_MASK": "1"}, {"TPLX_MASK_STAGE": "1"}):
    for k in ("TPLX_NO_MASK", "TPLX_MASK_STAGE"): os.environ.pop(k, None)
    os.environ.update(env)
    res = backend.Stage(prog).run_host(0, cols, n, seed)
    e = res.exceptions()
    rows = e["row"]
    print(env, "n_exc", len(e), "sorted", bool(np.all(np.diff(rows) > 0)), "first", rows[:5].tolist(), "last", rows[-3:].tolist(),
          "neq_idx", np.nonzero(rows != ora.exceptions["row"])[0][:5].tolist() if len(e) == len(ora.exceptions) else None)
    bad = np.nonzero(rows != ora.exceptions["row"])[0] if len(e) == len(ora.exceptions) else []
    if len(bad): i = int(bad[0]); print("   at", i, "gpu", e[max(0,i-2):i+3], "ora", ora.exceptions[max(0,i-2):i+3])
