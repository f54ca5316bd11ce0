"""The oracle (oracle/tplx_oracle.c + workloads.c) pinned against the reference's own golden vectors:
Zillow md5 (from the reference's zillow.cpp / runpython.py), TPC-H Q6 gtest golden, arithmetic semantics."""
import hashlib

import numpy as np

from tuplex_b200 import frontend, ir, workloads
from tuplex_b200.backend import Column
from tuplex_b200.ir import T_F64, T_I64, T_STR
from oracle import pyoracle


def test_zillow_pipeline_md5(built):
    cols, n = workloads.load_zillow_fixture()
    assert n == 32661
    prog = workloads.zillow_program()
    res = pyoracle.run_program(prog, cols, n)
    assert res.n_out == 577 and len(res.exceptions) == 0
    txt = workloads.rows_to_csv([res.values(c) for c in range(len(res.columns))], workloads.ZILLOW_OUT)
    assert hashlib.md5(txt).hexdigest() == "4d5ca0263b1a5058341a369116dee83a"
    assert txt == workloads.zillow_golden_csv()
    # selectivities quoted in SURVEY.md 8d: 2.43 % pass type == 'house', 1.77 % reach the output
    pre = pyoracle.run_program(prog.prefilter, cols, n)
    assert abs(pre.n_out / n - 0.0243) < 0.002 and abs(res.n_out / n - 0.0177) < 0.001


def test_q6_gtest_golden(built):
    cols = workloads.load_lineitem_fixture()
    n = len(cols[0].data)
    assert n == 60175
    res = pyoracle.run_program(workloads.q6_program(), cols, n)
    assert repr(ir.bits_f64(res.acc_seq[0])) == "1193053.2252999984"          # tuplex/test/core/TPCH.cc:85-97
    assert abs(ir.bits_f64(res.acc_tree[0]) - 1193053.2252999984) <= 1e-4     # same tolerance as the gtest
    assert pyoracle.q6(*[c.data for c in cols]) == ir.bits_f64(res.acc_seq[0])
    # partition-parallel port combines partials in partition order: still within the gtest tolerance
    assert abs(pyoracle.q6(*[c.data for c in cols], part_rows=4096, threads=4) - 1193053.2252999984) <= 1e-4
    # count aggregate == number of lines (TPCH.cc:74-79)
    sc = frontend.StageCompiler(workloads.Q6_TYPES, workloads.Q6_COLS)
    cnt = sc.finish_aggregate(lambda a, x: a + 1, lambda a, b: a + b, 0, 100001)
    assert pyoracle.run_program(cnt, cols, n).acc_seq[0] == n


def test_constant_folding_quirk(built):
    """0.06 + 0.01 folds to exactly 0.07 in the reference (ReduceExpressionsVisitor.cc:257-334); only the folded
    bounds reproduce the golden on the non-preprocessed query text (benchmarks/tpch/Q06/runtuplex.py:104-107)."""
    cols = workloads.load_lineitem_fixture()
    n = len(cols[0].data)
    sc = frontend.StageCompiler(workloads.Q6_TYPES, workloads.Q6_COLS)
    sc.add_filter(lambda x: 19940101 <= x['l_shipdate'] < 19940101 + 10000, 100001)
    sc.add_filter(lambda x: 0.06 - 0.01 <= x['l_discount'] <= 0.06 + 0.01, 100002)
    sc.add_filter(lambda x: x['l_quantity'] < 24, 100003)
    prog = sc.finish_aggregate(lambda a, x: a + x[1] * x[2], lambda a, b: a + b, 0.0, 100004)
    res = pyoracle.run_program(prog, cols, n)
    assert repr(ir.bits_f64(res.acc_seq[0])) == "1193053.2252999984"
    assert 0.06 + 0.01 != 0.07  # CPython would have used 0.06999999999999999


def test_python_arithmetic_semantics(built):
    L = pyoracle.lib()
    for a in range(-25, 26):
        for b in [-7, -3, -1, 1, 2, 5]:
            assert L.tplx_o_floordiv(a, b) == a // b
            assert L.tplx_o_floormod(a, b) == a % b
    # README example semantics on the normal case: [1, 2, 4] -> (x, x*x)
    sc = frontend.StageCompiler([T_I64], [None])
    sc.add_map(lambda x: (x, x * x), 100001)
    res = pyoracle.run_program(sc.finish_memory(), [Column(T_I64, np.array([1, 2, 4]))], 3)
    assert list(zip(res.values(0), res.values(1))) == [(1, 1), (2, 4), (4, 16)]
    # python/tests/test_filter.py: [1..5].map(x*x).filter(x>10) == [16,25]; filter(2<x<=4) == [3,4]
    sc = frontend.StageCompiler([T_I64], [None])
    sc.add_map(lambda x: x * x, 100001)
    sc.add_filter(lambda x: x > 10, 100002)
    assert pyoracle.run_program(sc.finish_memory(), [Column(T_I64, np.arange(1, 6))], 5).values(0) == [16, 25]
    sc = frontend.StageCompiler([T_I64], [None])
    sc.add_filter(lambda x: 2 < x <= 4, 100001)
    assert pyoracle.run_program(sc.finish_memory(), [Column(T_I64, np.arange(1, 6))], 5).values(0) == [3, 4]
    # DataSetCollect.cc:218-246: x[0]/x[1] with zero divisors -> exceptions, rest 42.0
    sc = frontend.StageCompiler([T_I64, T_I64], [None, None])
    sc.add_map(lambda x: x[0] / x[1], 100001)
    r = pyoracle.run_program(sc.finish_memory(), [Column(T_I64, np.array([84, 1, 3])), Column(T_I64, np.array([2, 0, 0]))], 3)
    assert r.values(0) == [42.0] and r.exceptions["code"].tolist() == [136, 136] and r.exceptions["row_no"].tolist() == [1, 2]


def test_aggregate_by_key_goldens(built):
    """tuplex/test/core/AggregateTest.cc:249-275: ('abc',-3), ('xyz',10); tuple aggregate ('abc',4,-1), ('xyz',6,3)."""
    rows = [(1, "abc", 0), (2, "xyz", 1), (4, "xyz", 2), (3, "abc", -1)]
    cols = [Column.from_values([r[0] for r in rows], T_I64), Column.from_values([r[1] for r in rows], T_STR),
            Column.from_values([r[2] for r in rows], T_I64)]
    sc = frontend.StageCompiler([T_I64, T_STR, T_I64], ["col0", "col1", "col2"])
    p = sc.finish_hash(["col1"], lambda a, x: a + x[0] * x[2], lambda a, b: a + b, 0, 100001)
    r = pyoracle.run_program(p, cols, 4)
    assert sorted(zip(r.values(0), r.values(1))) == [("abc", -3), ("xyz", 10)]
    sc = frontend.StageCompiler([T_I64, T_STR, T_I64], ["col0", "col1", "col2"])
    p = sc.finish_hash(["col1"], lambda a, x: (a[0] + x[0], a[1] + x[2]), lambda a, b: (a[0] + b[0], a[1] + b[1]), (0, 0), 100001)
    r = pyoracle.run_program(p, cols, 4)
    assert sorted(zip(r.values(0), r.values(1), r.values(2))) == [("abc", 4, -1), ("xyz", 6, 3)]


def test_row_format_bytes(built):
    """Serializer layout (utils/src/Serializer.cc:1016-1117): slots, var-len info = offset | size<<32, NUL-terminated."""
    cols = [Column.from_values([7, -1], T_I64), Column.from_values(["ab", ""], T_STR), Column.from_values([1.5, 2.0], T_F64)]
    (part,) = pyoracle.to_partitions(cols, 2, 1 << 20)
    assert int.from_bytes(part[:8], "little") == 2
    row0 = part[8:]
    assert int.from_bytes(row0[0:8], "little", signed=True) == 7
    info = int.from_bytes(row0[8:16], "little")
    assert info >> 32 == 3 and (info & 0xFFFFFFFF) == 24  # 'ab\0'; payload starts 24 bytes after slot 1
    assert int.from_bytes(row0[24:32], "little") == 3       # total var-len bytes
    assert row0[32:35] == b"ab\0"
    # exception record (IExceptionableTask.h:22-36)
    exc = np.array([(1, 5, 136, 100001)], dtype=[("row", "<i8"), ("row_no", "<i8"), ("code", "<i8"), ("op_id", "<i8")])
    e = pyoracle.exception_partition(cols, exc)
    assert int.from_bytes(e[:8], "little") == 1
    assert [int.from_bytes(e[8 + 8 * k:16 + 8 * k], "little") for k in range(4)] == [5, 136, 100001, 33]
