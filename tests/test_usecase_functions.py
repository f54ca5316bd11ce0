"""The reference's UseCaseFunctionsTest.cc known answers (tests/usecase_cases.py) through the UDF front end and the CPU oracle; the GPU
twin runs the same vectors through Context. Cases the front end does not lower take the interpreter path in the product: here they are
checked against the gtest golden with CPython and listed."""
import math

import pytest

import tuplex_b200
from oracle import pyoracle
from tuplex_b200 import frontend, pyexec
from usecase_cases import CASES, COLUMN_CASES, OPTION_CASES


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return type(a) is type(b) and a == b


def _cpython(rows, udf):
    fn = pyexec.udf_from_source(udf)
    out = []
    for r in rows:
        out.append(fn(*r) if isinstance(r, tuple) and fn.__code__.co_argcount == len(r) and len(r) > 1 else fn(r))
    return out


def test_usecase_vectors_lower_and_match_goldens(built):
    ctx = tuplex_b200.Context()
    lowered, not_lowered = 0, []
    for name, line, rows, udf, expected in CASES:
        got_py = _cpython(rows, udf)
        assert all(_same(g, w) for g, w in zip(got_py, expected)) and len(got_py) == len(expected), (name, got_py, expected)  # the golden is Python's answer
        src = ctx._source_from_rows(list(rows), None)
        assert not src.fallback, name
        sc = frontend.StageCompiler([c.type for c in src.cols], src.names)
        try:
            sc.add_map(udf, 100001)
            prog = sc.finish_memory()
        except frontend.UnsupportedUDF as e:
            not_lowered.append((name, str(e)))
            continue
        lowered += 1
        res = pyoracle.run_program(prog, src.cols, src.n_rows)
        assert not len(res.exceptions), name
        cols = [res.values(c) for c in range(len(res.columns))]
        got = [tuple(col[i] for col in cols) if len(cols) != 1 else cols[0][i] for i in range(res.n_out)]
        assert len(got) == len(expected), (name, got, expected)
        for g, w in zip(got, expected):
            assert _same(g, w), (f"{name} (UseCaseFunctionsTest.cc:{line})", g, w)
    # str() of floats / bools and strip(chars) are outside the device op set: interpreter path
    assert {n for n, _ in not_lowered} <= {"StrCast.bool", "StrCast.float", "strFormatFunction.v5", "strStripChars"}, not_lowered
    assert lowered >= 29


def test_usecase_option_vectors_match_cpython(built):
    """NestedIf / FloatNullError: UDFs over Option inputs (the gtest only prints; CPython decides), lowered with the None-aware front end."""
    import numpy as np
    from tuplex_b200.backend import Column
    from tuplex_b200.dataset import _option_cols
    from tuplex_b200.ir import T_BOOL
    ctx = tuplex_b200.Context()
    for name, line, rows, names, kind, col, udf in OPTION_CASES:
        src = ctx._source_from_rows(list(rows), names)
        assert not src.fallback
        sc = frontend.StageCompiler([c.type for c in src.cols], src.names, _option_cols(src.cols))
        try:
            (sc.add_with_column(col, udf, 100001) if kind == "withColumn" else sc.add_map(udf, 100001))
            prog = sc.finish_memory()
        except frontend.UnsupportedUDF:
            assert name == "FloatNullError"  # a column that is None for every row: NULLVALUE type, interpreter path
            continue
        cols = list(src.cols)
        for j in sorted(prog.null_of):
            pres = src.cols[prog.null_of[j]].present()
            cols.append(Column(T_BOOL, (~pres).astype(np.int64) if pres is not None else np.zeros(src.n_rows, np.int64)))
        res = pyoracle.run_program(prog, cols, src.n_rows)
        ops = [pyexec.Op(kind, udf, column=col)]
        want, raised = [], []
        for i, r in enumerate(rows):
            try:
                want.append(pyexec.run_row(ops, r, names or [None])[0])
            except TypeError:
                raised.append(i)  # float(None): the row leaves the device as a TypeError exception row
        assert [int(e["row"]) for e in res.exceptions] == raised and all(int(e["code"]) == 129 for e in res.exceptions), name
        vis = len(prog.out_cols) - prog.hidden_out_cols
        vals = [res.values(k) for k in range(len(res.columns))]
        for k, nof in enumerate(prog.out_null_of):
            if nof:
                vals[nof - 1] = [None if isnull else v for v, isnull in zip(vals[nof - 1], vals[k])]
        got = list(zip(*vals[:vis]))
        assert got == [w if isinstance(w, tuple) else (w,) for w in want], (name, got, want)


@pytest.mark.gpu
def test_usecase_vectors_through_context(gpu):
    c = tuplex_b200.Context()
    for name, line, rows, udf, expected in CASES:
        got = c.parallelize(list(rows)).map(udf).collect()
        assert len(got) == len(expected) and all(_same(g, w) for g, w in zip(got, expected)), (f"{name} (UseCaseFunctionsTest.cc:{line})", got, expected)
    for name, line, rows, names, kind, col, udf, expected in COLUMN_CASES:
        ds = c.parallelize(list(rows), columns=names) if names else c.parallelize(list(rows))
        ds = ds.map(udf) if kind == "map" else ds.filter(udf) if kind == "filter" else ds.withColumn(col, udf)
        got = ds.collect()
        assert len(got) == len(expected) and all(_same(g, w) for g, w in zip(got, expected)), (name, got, expected)
    for name, line, rows, names, kind, col, udf in OPTION_CASES:
        ds = c.parallelize(list(rows), columns=names) if names else c.parallelize(list(rows))
        got = (ds.withColumn(col, udf) if kind == "withColumn" else ds.map(udf)).collect()
        ops = [pyexec.Op(kind, udf, column=col)]
        want = []
        for r in rows:
            try:
                want.append(pyexec.run_row(ops, r, names or [None])[0])
            except TypeError:
                pass  # stays an exception row (no resolver)
        assert got == want, name
