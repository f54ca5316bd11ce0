"""`**`: front end + oracle vs CPython for integer bases (exact) and vs the reference's multiply chains for float bases."""
import math

import pytest

from tuplex_b200 import frontend
from tuplex_b200.ir import T_F64, T_I64
from oracle import pyoracle
import power_udfs as U


def _prog(src):
    sc = frontend.StageCompiler([T_I64, T_F64], ["a", "f"])
    sc.add_with_column("r", src, 100001)
    sc.add_select(["r"], 100002)
    return sc.finish_memory()


def _chain(b, k):
    n = abs(k)
    if n == 0:
        return 1.0
    if b == 0.0 and k > 0:
        return 0.0
    b2 = b * b
    p = {1: b, 2: b2, 3: b2 * b, 4: b2 * b2, 5: b * (b2 * b2), 6: b2 * (b2 * b2)}[n]
    return 1.0 / p if k < 0 else p


@pytest.mark.parametrize("src,k", U.INT_CASES + U.NEG_CASES)
def test_integer_power_matches_cpython(built, src, k):
    cols, rows = U.make_columns(2000, 1)
    res = pyoracle.run_program(_prog(src), cols, 2000)
    exc = {int(e["row"]): int(e["code"]) for e in res.exceptions}
    vals = res.values(0)
    j = 0
    for i, (a, f) in enumerate(rows):
        if a == 0 and k < 0:
            assert exc.get(i) == 136, (src, a)  # ZeroDivisionError
            continue
        assert i not in exc
        assert vals[j] == a ** k, (src, a, vals[j])
        j += 1


@pytest.mark.parametrize("src", U.MIXED)
def test_power_in_expressions(built, src):
    cols, rows = U.make_columns(1500, 2)
    res = pyoracle.run_program(_prog(src), cols, 1500)
    exc = {int(e["row"]) for e in res.exceptions}
    vals = res.values(0)
    fn = eval(src)
    j = 0
    for i, (a, f) in enumerate(rows):
        try:
            want = fn({"a": a, "f": f})
        except ZeroDivisionError:
            assert i in exc
            continue
        assert i not in exc and vals[j] == want, (src, a, vals[j], want)
        j += 1


@pytest.mark.parametrize("src,k", U.FLOAT_CASES)
def test_float_power_is_the_reference_multiply_chain(built, src, k):
    cols, rows = U.make_columns(2000, 3)
    res = pyoracle.run_program(_prog(src), cols, 2000)
    exc = {int(e["row"]) for e in res.exceptions}
    vals = res.values(0)
    j = 0
    for i, (a, f) in enumerate(rows):
        if f == 0.0 and k < 0:
            assert i in exc
            continue
        want = _chain(f, k)
        assert i not in exc and (vals[j] == want and math.copysign(1, vals[j]) == math.copysign(1, want)), (src, f, vals[j], want)
        # (CPython's own f ** 2 goes through libm pow and can differ from f * f in the last bit; the reference multiplies)
        j += 1


@pytest.mark.parametrize("src", U.UNSUPPORTED)
def test_other_exponents_stay_on_the_interpreter_path(src):
    with pytest.raises(frontend.UnsupportedUDF):
        _prog(src)
