"""The reference's Python-level tests (tuplex/python/tests/test_strings.py, test_arithmetic.py, test_columns.py, …) restated
as vectors in tests/ref_python_cases.py, run through tuplex_b200.Context on the GPU exactly as the reference tests run them
through tuplex.Context: parallelize(rows).map(udf).collect() == expected."""
import math

import pytest

import tuplex_b200
from ref_python_cases import COLUMN_CASES, MAP_CASES, expected_of

pytestmark = pytest.mark.gpu


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


def test_map_vectors(gpu):
    ctx = tuplex_b200.Context()
    for name, rows, udf, expected in MAP_CASES:
        want = expected_of(rows, udf, expected)
        got = ctx.parallelize(list(rows)).map(udf).collect()
        assert len(got) == len(want) and all(_same(g, w) for g, w in zip(got, want)), (name, got, want)
    # every one of these UDFs is inside the GPU op set: none took the whole-stage CPython path
    assert not [m for m in ctx._messages if "falls back" in m], ctx._messages


def test_column_vectors(gpu):
    ctx = tuplex_b200.Context()
    for name, rows, columns, build, want in COLUMN_CASES:
        got = build(ctx.parallelize(list(rows), columns=columns)).collect()
        assert len(got) == len(want) and all(_same(g, w) for g, w in zip(got, want)), (name, got, want)


def test_rename_column(gpu):
    """test_columns.py:89-99"""
    ctx = tuplex_b200.Context()
    ds = ctx.parallelize([(1, 2), (3, 2)])
    ds2 = ds.renameColumn(0, "first")
    assert ds2.columns[0] == "first"
    ds3 = ds2.renameColumn(1, "second")
    assert ds3.columns[1] == "second"
    ds4 = ds3.renameColumn("first", "1")
    assert ds4.columns == ["1", "second"]
