"""float(str) (TPLX_OP_S2F): the reference casts with fast_atod behind the runtime's trimming wrapper
(tuplex/runtime/src/Runtime.cc:343-365, utils/src/StringUtils.cc:71-163) — not correctly rounded, prefix quirks of the
nan / infinity matcher, ValueError otherwise. Front end + oracle here; the GPU twin is in tests/test_gpu_parity.py."""
import math
import random
import struct

from tuplex_b200 import frontend
from tuplex_b200.backend import Column
from tuplex_b200.ir import T_STR
from oracle import pyoracle as po


def float_cases():
    rng = random.Random(31)
    vals = ["1.5", " 2.25\t", "-0.07", "1801.0", "1e5", "1e-2", "nan", "NaN", "inf", "Infinity", "-inf", "+3.5", ".5", "5.", "", "  ", "abc", "1.2.3",
            "n", "infi", "12a", "0.000000005", "123456789.123456789", "1e400"]
    vals += [f"{rng.uniform(-1e6, 1e6):.{rng.randint(0, 12)}f}" for _ in range(3000)]
    vals += [str(rng.randint(-10**12, 10**12)) for _ in range(500)]
    return vals


def test_float_of_str_matches_fast_atod(built):
    vals = float_cases()
    sc = frontend.StageCompiler([T_STR], ["s"])
    sc.add_map(lambda x: float(x["s"]) * 2.0, 100001)
    prog = sc.finish_memory()
    col = Column.from_values(vals, T_STR)
    res = po.run_program(prog, [col], len(vals))
    want = [po.csv_scalar("f64", v) for v in vals]
    exc_rows = {int(e["row"]): int(e["code"]) for e in res.exceptions}
    got = list(res.values(0))
    k = 0
    for i, w in enumerate(want):
        if w is None:
            assert exc_rows.get(i) == 135, (vals[i], exc_rows.get(i))  # ValueError
        else:
            assert i not in exc_rows, vals[i]
            g = got[k]
            k += 1
            assert struct.pack("<d", g) == struct.pack("<d", w * 2.0) or (math.isnan(g) and math.isnan(w)), (vals[i], g, w)
    assert k == res.n_out
