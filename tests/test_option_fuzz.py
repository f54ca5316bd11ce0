"""Random None-aware UDFs over Option[T] columns: front end + oracle (companion columns expanded like the executor does) against CPython
evaluating the same lambda source, row by row — values, None results, dropped rows and TypeError rows (a None that is used)."""
import random

import numpy as np
import pytest

from oracle import pyoracle
from tuplex_b200 import frontend
from tuplex_b200.backend import Column
from tuplex_b200.ir import C, T_BOOL, T_F64, T_I64, T_STR
from tuplex_b200.pyexec import Row

NAMES = ["a", "b", "s", "f", "t"]
TYPES = [T_I64, T_I64, T_STR, T_F64, T_STR]
OPTION = [0, 2, 3]
WORDS = ["", "a", "abc", "House", "x y", "12"]


class Gen:
    def __init__(self, seed):
        self.r = random.Random(seed)

    def oint(self, d=1):  # Option[int] expression (may be None)
        r = self.r
        k = r.randint(0, 4) if d > 0 else 0
        if k <= 1:
            return "x['a']"
        if k == 2:
            return f"(None if {self.bool_(0)} else {self.int_(0)})"
        if k == 3:
            return f"(x['a'] if {self.bool_(0)} else None)"
        return f"(x['a'] if x['b'] > {r.randint(-3, 3)} else x['b'])"

    def ostr(self, d=1):
        r = self.r
        k = r.randint(0, 3) if d > 0 else 0
        if k <= 1:
            return "x['s']"
        if k == 2:
            return f"(None if {self.bool_(0)} else x['t'])"
        return f"(x['s'] if {self.bool_(0)} else x['t'].upper())"

    def int_(self, d):  # int expression that is never None (may raise TypeError where it uses a None)
        r = self.r
        k = r.random()
        if d <= 0 or k < 0.25:
            return r.choice(["x['b']", str(r.randint(-5, 5)), "len(x['t'])"])
        if k < 0.45:
            return f"({self.int_(d - 1)} {r.choice(['+', '-', '*'])} {self.int_(d - 1)})"
        if k < 0.6:
            o = self.oint()
            return f"({self.int_(d - 1)} if {o} is None else {o} + {r.randint(0, 3)})"
        if k < 0.7:
            return f"(x['a'] if x['a'] is not None else {self.int_(d - 1)})"
        if k < 0.78:
            return f"(len(x['s']) if x['s'] is not None else {r.randint(-2, 2)})"
        if k < 0.86:
            return f"(x['a'] {r.choice(['+', '*', '-'])} {self.int_(d - 1)})"        # uses a possibly-None value: TypeError rows
        if k < 0.93:
            return f"({self.int_(d - 1)} if {self.bool_(d - 1)} else {self.int_(d - 1)})"
        return f"(x['a'] if x['a'] else {r.randint(1, 4)})"                          # None and 0 are both falsy

    def bool_(self, d):
        r = self.r
        k = r.random()
        if d <= 0 or k < 0.2:
            return r.choice(["x['a'] is None", "x['s'] is not None", "x['f'] is None", f"x['b'] > {r.randint(-3, 3)}", "x['a'] == None", "x['s'] != None"])
        if k < 0.35:
            return f"({self.oint()} {r.choice(['==', '!='])} {self.int_(d - 1)})"
        if k < 0.45:
            return f"({self.oint()} {r.choice(['==', '!='])} {self.oint()})"
        if k < 0.55:
            return f"({self.ostr()} {r.choice(['==', '!='])} {r.choice(WORDS)!r})"
        if k < 0.65:
            return f"(x['a'] is not None and x['a'] {r.choice(['<', '>', '<=', '>='])} {self.int_(d - 1)})"
        if k < 0.75:
            return f"(x['s'] is None or {r.choice(WORDS)!r} in x['s'])"
        if k < 0.82:
            return f"bool({r.choice([self.oint, self.ostr, lambda: 'x[' + repr('f') + ']'])()})"
        if k < 0.9:
            return f"(not {self.bool_(d - 1)})"
        if k < 0.95:
            return f"({self.bool_(d - 1)} {r.choice(['and', 'or'])} {self.bool_(d - 1)})"
        return f"(x['f'] is not None and x['f'] > {r.choice(['0.5', '-1.0', '2.25'])})"

    def out(self):
        r = self.r
        makers = [self.oint, lambda: self.int_(2), self.ostr, lambda: self.bool_(2), lambda: "x['f']", lambda: f"(x['f'] if {self.bool_(1)} else None)",
                  lambda: "(x['s'].lower() if x['s'] else 'none')", lambda: "(None if x['a'] is None else x['a'] * 2)"]
        elems = [r.choice(makers)() for _ in range(r.randint(1, 4))]
        return "(" + ", ".join(elems) + ",)"


def _data(n, seed):
    rng = random.Random(seed)
    a = [None if rng.random() < 0.3 else rng.randint(-3, 6) for _ in range(n)]
    b = [rng.randint(-4, 4) for _ in range(n)]
    s = [None if rng.random() < 0.25 else rng.choice(WORDS) for _ in range(n)]
    f = [None if rng.random() < 0.25 else rng.choice([0.0, 0.5, -1.0, 2.25, 7.5]) for _ in range(n)]
    t = [rng.choice(WORDS) for _ in range(n)]
    return [a, b, s, f, t]


def _expanded(cols, prog):
    out = list(cols)
    for j in sorted(prog.null_of):
        pres = cols[prog.null_of[j]].present()
        out.append(Column(T_BOOL, (~pres).astype(np.int64) if pres is not None else np.zeros(len(cols[0]), np.int64)))
    return out


def _rows(prog, res):
    n_vis = len(prog.out_cols) - prog.hidden_out_cols
    cols = [res.values(k) for k in range(len(res.columns))]
    for k, nof in enumerate(prog.out_null_of):
        if nof and k < len(cols):
            cols[nof - 1] = [None if isnull else v for v, isnull in zip(cols[nof - 1], cols[k])]
    return list(zip(*cols[:n_vis]))


@pytest.mark.parametrize("seed", range(6))
def test_random_option_udfs_match_cpython(built, seed):
    g = Gen(900 + seed)
    n = 400
    vals = _data(n, seed)
    cols = [Column.from_values(v, t) for v, t in zip(vals, TYPES)]
    compared = 0
    for trial in range(40):
        flt, mp = "lambda x: " + g.bool_(2), "lambda x: " + g.out()
        sc = frontend.StageCompiler(TYPES, NAMES, OPTION)
        try:
            sc.add_filter(flt, 100001)
            sc.add_map(mp, 100002)
            prog = sc.finish_memory()
        except frontend.UnsupportedUDF:
            continue  # e.g. a column that is None for every row
        res = pyoracle.run_program(prog, _expanded(cols, prog), n)
        got = _rows(prog, res)
        exc = {int(e["row"]): int(e["code"]) for e in res.exceptions}
        ffn, mfn = eval(flt), eval(mp)
        j = 0
        for i in range(n):
            row = Row([v[i] for v in vals], NAMES)
            try:
                keep = ffn(row)
                want = mfn(row) if keep else None
            except TypeError:
                assert exc.get(i) == C["TPLX_EC_TYPEERROR"], (flt, mp, list(row), exc.get(i))
                continue
            assert i not in exc, (flt, mp, list(row), exc.get(i))
            if not keep:
                continue
            assert j < len(got) and got[j] == want, (flt, mp, list(row), got[j] if j < len(got) else None, want)
            j += 1
        assert j == len(got), (flt, mp)
        compared += 1
    assert compared >= 25
