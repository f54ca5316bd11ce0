"""The reference's Python-level test vectors (tests/ref_python_cases.py) through the UDF front end and the CPU oracle.
GPU twin: tests/test_gpu_reference_pytests.py (same vectors through Context / the C ABI)."""
import math

import numpy as np
import pytest

import tuplex_b200
from tuplex_b200 import frontend
from oracle import pyoracle
from ref_python_cases import MAP_CASES, expected_of


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b and type(a) == type(b) or (isinstance(a, (int, float)) and isinstance(b, (int, float)) and a == b and
                                            isinstance(a, bool) == isinstance(b, bool))


def test_reference_python_vectors_lower_and_match(built):
    ctx = tuplex_b200.Context()
    lowered = 0
    not_lowered = []
    for name, rows, udf, expected in MAP_CASES:
        want = expected_of(rows, udf, expected)
        src = ctx._source_from_rows(list(rows), None, infer=True)
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
        cols = [res.values(c) for c in range(len(res.columns))]
        got = [tuple(col[i] for col in cols) if len(cols) != 1 else cols[0][i] for i in range(res.n_out)]
        # bool outputs come back as 0/1 slots typed bool by the program
        assert len(got) == len(want), (name, got, want)
        for g, w in zip(got, want):
            assert _same(g, w), (name, g, w)
    assert lowered >= 30, not_lowered
