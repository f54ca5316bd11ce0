"""K1v planner (vec_plan in tplx_gpu.cu) on the host: the micro-op program it emits — accumulator chains, fused compare / filter,
dropped stores, renumbered slots — run by tests/vec_emul.py must give exactly what the oracle gives for the op program itself:
same kept rows (bit-exact values), same exception rows, codes and operators. Needs no GPU (tplx_gpu_stage_vec_plan)."""
import random

import numpy as np
import pytest

from tuplex_b200 import backend, frontend, workloads
from tuplex_b200.backend import Column
from tuplex_b200.ir import T_F64, T_I64
from oracle import pyoracle
import vec_emul as E

COLS = ["a", "b", "f", "g"]
TYPES = [T_I64, T_I64, T_F64, T_F64]


class NumGen:
    """random numeric UDFs (no strings): what K1v is eligible for"""

    def __init__(self, seed):
        self.r = random.Random(seed)

    def int_(self, d):
        r = self.r
        if d <= 0 or r.random() < 0.25:
            return r.choice(["x['a']", "x['b']", str(r.randint(-9, 9)), "x['a']"])
        k = r.random()
        if k < 0.5:
            op = r.choice(["+", "-", "*", "//", "%", "&", "|", "^", "% 4", "// 8", "% 2", "<< 2", ">> 1"])
            if " " in op:
                return f"({self.int_(d - 1)} {op})"
            return f"({self.int_(d - 1)} {op} {self.int_(d - 1)})"
        if k < 0.6:
            return f"(-{self.int_(d - 1)})"
        if k < 0.7:
            return f"abs({self.int_(d - 1)})"
        if k < 0.85:
            return f"({self.int_(d - 1)} if {self.bool_(d - 1)} else {self.int_(d - 1)})"
        if k < 0.93:
            return f"{r.choice(['min', 'max'])}({self.int_(d - 1)}, {self.int_(d - 1)})"
        return f"int({self.float_(d - 1)})"

    def float_(self, d):
        r = self.r
        if d <= 0 or r.random() < 0.3:
            return r.choice(["x['f']", "x['g']", "2.5", "0.1", "-3.0"])
        k = r.random()
        if k < 0.5:
            return f"({self.float_(d - 1)} {r.choice(['+', '-', '*', '/', '%'])} {self.float_(d - 1)})"
        if k < 0.7:
            return f"({self.int_(d - 1)} {r.choice(['+', '*', '/'])} {self.float_(d - 1)})"
        if k < 0.8:
            return f"({self.int_(d - 1)} / {self.int_(d - 1)})"
        if k < 0.9:
            return f"abs({self.float_(d - 1)})"
        return f"float({self.int_(d - 1)})"

    def bool_(self, d):
        r = self.r
        k = r.random()
        if d <= 0 or k < 0.4:
            return f"({self.int_(max(d, 0))} {r.choice(['<', '<=', '>', '>=', '==', '!='])} {self.int_(0)})"
        if k < 0.6:
            return f"({self.float_(d - 1)} {r.choice(['<', '>=', '==', '!='])} {self.float_(d - 1)})"
        if k < 0.85:
            return f"({self.bool_(d - 1)} {r.choice(['and', 'or'])} {self.bool_(d - 1)})"
        return f"(not {self.bool_(d - 1)})"

    def pipeline(self):
        r = self.r
        ops, names = [], []
        for i in range(r.randint(1, 4)):
            kind = r.choice(["int_", "int_", "float_", "bool_"])
            ops.append(("col", f"c{i}", f"lambda x: {getattr(self, kind)(r.randint(1, 3))}"))
            names.append(f"c{i}")
            if r.random() < 0.5:
                ops.append(("filter", f"lambda x: {self.bool_(r.randint(0, 2))}"))
        if r.random() < 0.6:
            ops.append(("select", r.sample(names, r.randint(1, len(names))) + r.sample(COLS, r.randint(0, 2))))
        return ops


def make_columns(n, seed):
    rnd = np.random.default_rng(seed)
    a = rnd.integers(-50, 50, n).astype(np.int64)
    b = rnd.integers(-4, 5, n).astype(np.int64)
    f = np.round(rnd.normal(0, 20, n), 2)
    f[rnd.integers(0, n, max(1, n // 20))] = 0.0
    g = np.round(rnd.normal(1, 3, n), 1)
    return [Column(T_I64, a), Column(T_I64, b), Column(T_F64, f), Column(T_F64, g)]


def compile_ops(ops):
    sc = frontend.StageCompiler(TYPES, COLS)
    k = 100001
    for op in ops:
        if op[0] == "col":
            sc.add_with_column(op[1], op[2], k)
        elif op[0] == "filter":
            sc.add_filter(op[1], k)
        else:
            sc.add_select(op[1], k)
        k += 1
    return sc.finish_memory(prefilter=False)


def test_c1_plan_is_three_micro_ops():
    """config 0 (x*x, x % 2 == 0): load -> multiply (the one stored value: the output column) -> masked compare that filters"""
    st = backend.Stage(workloads.c1_program())
    uops, n_slots, out_slots = st.vec_plan()
    st.close()
    assert [E.NAME[u["vop"]] for u in uops] == ["V_LDCOL", "V_IMUL", "V_ICMP_EQ"]
    assert uops[0]["xflags"] == E.X_NOSTORE
    assert uops[1]["xflags"] == E.X_A_ACC | E.X_B_ACC and uops[1]["dst"] == out_slots[0]
    assert uops[2]["xflags"] == E.X_A_ACC | E.X_NOSTORE | E.X_FILTER | E.X_A_MASK and uops[2]["imm2"] == 1
    assert n_slots == 1


@pytest.mark.parametrize("seed", range(40))
def test_planned_micro_ops_equal_the_op_program(seed):
    g = NumGen(7000 + seed)
    n = 400
    cols = make_columns(n, seed)
    raw = [np.ascontiguousarray(c.data).view(np.uint64).tolist() for c in cols]
    compared = 0
    for trial in range(12):
        ops = g.pipeline()
        try:
            prog = compile_ops(ops)
        except frontend.UnsupportedUDF:
            continue
        st = backend.Stage(prog)
        plan = st.vec_plan()
        st.close()
        if plan is None:
            continue
        uops, n_slots, out_slots = plan
        used = {s for u in uops for s in (u["dst"], u["a"], u["b"], u["c"], u["guard"]) if s != E.NOSLOT} | set(out_slots)
        assert all(s < n_slots for s in used), ops
        got_rows, got_exc = E.run(uops, out_slots, raw, n)
        ora = pyoracle.run_program(prog, cols, n, 0)
        exp_cols = [np.ascontiguousarray(ora.columns[c][1]).view(np.uint64)[:ora.n_out].tolist() for c in range(len(out_slots))]
        exp_rows = list(zip(*exp_cols)) if exp_cols else []
        assert len(got_rows) == ora.n_out, f"seed {seed} trial {trial}: {ops}"
        assert got_rows == exp_rows, f"seed {seed} trial {trial}: {ops}"
        exp_exc = [(int(e["row"]), int(e["code"]), int(e["op_id"])) for e in ora.exceptions]
        assert [(r, c, prog.opids[o]) for r, c, o in got_exc] == exp_exc, f"seed {seed} trial {trial}: {ops}"
        compared += 1
    assert compared >= 4
