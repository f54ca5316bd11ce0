"""Random, type-correct UDF generator over the GPU op set (shared by the CPU and GPU fuzz tests).
UDFs are produced as source strings (the front end accepts `lambda` source like the reference's C++ tests
pass UDF("lambda x: ...")), so every generated pipeline can also be evaluated by CPython with eval()."""
import random

import numpy as np

from tuplex_b200.backend import Column
from tuplex_b200.ir import T_F64, T_I64, T_STR

COLS = ["a", "b", "f", "s", "t"]
TYPES = [T_I64, T_I64, T_F64, T_STR, T_STR]
WORDS = ["", "a", "Ab", "house", "HOUSE for sale", "12", " 7", "-3", "x9", "3 bds , 2 ba", "a,b,,c", "  pad  ", "Zz", "condo/unit", "0042", "$1,250/mo"]
NEEDLES = ["a", "b", ",", " ", "ho", "HO", "12", "", "bd", "/", "xyz"]


def make_columns(n, seed):
    rnd = np.random.default_rng(seed)
    a = rnd.integers(-50, 50, n)
    b = rnd.integers(-4, 5, n)
    f = np.round(rnd.normal(0, 20, n), 2)
    f[rnd.integers(0, n, max(1, n // 20))] = 0.0
    s = [WORDS[i] for i in rnd.integers(0, len(WORDS), n)]
    t = [WORDS[i] + WORDS[j][:3] for i, j in zip(rnd.integers(0, len(WORDS), n), rnd.integers(0, len(WORDS), n))]
    cols = [Column(T_I64, a.astype(np.int64)), Column(T_I64, b.astype(np.int64)), Column(T_F64, f.astype(np.float64)),
            Column.from_values(s, T_STR), Column.from_values(t, T_STR)]
    rows = list(zip(a.tolist(), b.tolist(), f.tolist(), s, t))
    return cols, rows


class Gen:
    def __init__(self, seed):
        self.r = random.Random(seed)

    def int_(self, d):
        r = self.r
        if d <= 0 or r.random() < 0.25:
            return r.choice(["x['a']", "x['b']", str(r.randint(-9, 9)), "x['a']", "len(x['s'])"])
        k = r.random()
        if k < 0.45:
            op = r.choice(["+", "-", "*", "//", "%", "&", "|", "^"])
            return f"({self.int_(d - 1)} {op} {self.int_(d - 1)})"
        if k < 0.55:
            return f"(-{self.int_(d - 1)})"
        if k < 0.65:
            return f"abs({self.int_(d - 1)})"
        if k < 0.75:
            return f"{self.str_(d - 1)}.find({r.choice(NEEDLES)!r})"
        if k < 0.82:
            return f"{self.str_(d - 1)}.rfind({r.choice(NEEDLES)!r})"
        if k < 0.88:
            return f"int({self.str_(d - 1)})"
        if k < 0.92:
            return f"({self.int_(d - 1)} if {self.bool_(d - 1)} else {self.int_(d - 1)})"
        if k < 0.96:
            return f"{r.choice(['min', 'max'])}({self.int_(d - 1)}, {self.int_(d - 1)})"
        return f"int({self.float_(d - 1)})"

    def float_(self, d):
        r = self.r
        if d <= 0 or r.random() < 0.3:
            return r.choice(["x['f']", "x['f']", "2.5", "0.1", "-3.0"])
        k = r.random()
        if k < 0.5:
            op = r.choice(["+", "-", "*", "/", "%"])
            return f"({self.float_(d - 1)} {op} {self.float_(d - 1)})"
        if k < 0.75:
            return f"({self.int_(d - 1)} {r.choice(['+', '*', '/'])} {self.float_(d - 1)})"
        if k < 0.85:
            return f"({self.int_(d - 1)} / {self.int_(d - 1)})"
        return f"float({self.int_(d - 1)})"

    def str_(self, d):
        r = self.r
        if d <= 0 or r.random() < 0.3:
            return r.choice(["x['s']", "x['t']", "x['s']", repr(r.choice(WORDS))])
        k = r.random()
        if k < 0.2:
            return f"{self.str_(d - 1)}.{r.choice(['lower', 'upper', 'strip', 'lstrip', 'rstrip'])}()"
        if k < 0.4:
            lo = r.choice(["", str(r.randint(-3, 4)), self.int_(0)])
            hi = r.choice(["", str(r.randint(-3, 6)), self.int_(0)])
            return f"{self.str_(d - 1)}[{lo}:{hi}]"
        if k < 0.5:
            return f"{self.str_(d - 1)}[{r.randint(-2, 2)}]"
        if k < 0.65:
            return f"({self.str_(d - 1)} + {self.str_(d - 1)})"
        if k < 0.78:
            return f"{self.str_(d - 1)}.replace({r.choice(NEEDLES)!r}, {r.choice(['', '_', 'xy'])!r})"
        if k < 0.82:
            return f"({r.choice(['%d', '%05d', 'n=%d!', '%3d'])!r} % {self.int_(d - 1)})"
        if k < 0.86:
            return r.choice([f"{r.choice(['{:03}|{}', 'v={}', '{1}-{0:4d}'])!r}.format({self.int_(d - 1)}, {self.str_(0)})",
                             "f\"{x['a']:04}:{x['s']}\""])
        if k < 0.93:
            return f"({self.str_(d - 1)} if {self.bool_(d - 1)} else {self.str_(d - 1)})"
        return f"str({self.int_(d - 1)})"

    def bool_(self, d):
        r = self.r
        k = r.random()
        if d <= 0 or k < 0.3:
            op = r.choice(["<", "<=", ">", ">=", "==", "!="])
            return f"({self.int_(0)} {op} {self.int_(0)})"
        if k < 0.45:
            return f"({self.float_(d - 1)} {r.choice(['<', '>=', '==', '!='])} {self.float_(d - 1)})"
        if k < 0.55:
            return f"({r.choice(NEEDLES)!r} in {self.str_(d - 1)})"
        if k < 0.65:
            return f"({self.str_(d - 1)} {r.choice(['==', '!='])} {self.str_(d - 1)})"
        if k < 0.69:
            return f"{self.str_(d - 1)}.{r.choice(['startswith', 'endswith'])}({r.choice(NEEDLES)!r})"
        if k < 0.72:
            return r.choice([f"({self.int_(0)} {r.choice(['in', 'not in'])} (1, -2, 7))", f"({self.str_(0)} in ['a', 'house', ''])"])
        if k < 0.86:
            return f"({self.bool_(d - 1)} {r.choice(['and', 'or'])} {self.bool_(d - 1)})"
        if k < 0.93:
            return f"(not {self.bool_(d - 1)})"
        return f"({self.int_(d - 1)} < {self.int_(d - 1)} <= {self.int_(d - 1)})"

    def pipeline(self):
        """[(kind, arg...)] using the StageCompiler method names; at most one filter early, outputs of mixed types."""
        r = self.r
        ops = []
        n_new = r.randint(1, 3)
        for i in range(n_new):
            kind = r.choice(["int", "float", "str", "bool"])
            expr = getattr(self, {"int": "int_", "float": "float_", "str": "str_", "bool": "bool_"}[kind])(r.randint(1, 3))
            ops.append(("add_with_column", f"c{i}", f"lambda x: {expr}"))
            if r.random() < 0.4:
                ops.append(("add_filter", f"lambda x: {self.bool_(r.randint(0, 2))}"))
        if r.random() < 0.5:
            ops.append(("add_select", [f"c{i}" for i in range(n_new)] + r.sample(COLS, r.randint(0, 2))))
        return ops


def apply_ops(sc, ops, first_id=100001):
    k = first_id
    for op in ops:
        if op[0] == "add_with_column":
            sc.add_with_column(op[1], op[2], k)
        elif op[0] == "add_filter":
            sc.add_filter(op[1], k)
        else:
            sc.add_select(op[1], k)
        k += 1
