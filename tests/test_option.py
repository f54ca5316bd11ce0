"""Option[T] / None-aware normal case: front end + oracle against CPython for UDFs over columns that hold None, the
descriptor's companion columns, and (GPU) the same programs through the C ABI with validity bitmaps in and out."""
import random

import numpy as np
import pytest

from oracle import pyoracle
from tuplex_b200 import backend, frontend, ir
from tuplex_b200.backend import Column
from tuplex_b200.ir import C, T_BOOL, T_I64, T_STR
from tuplex_b200.pyexec import Row

from option_udfs import UDFS

NAMES = ["a", "b", "c", "s"]
TYPES = [T_I64, T_I64, T_I64, T_STR]
OPTION = [0, 2, 3]


def _data(n, seed):
    rng = random.Random(seed)
    a = [None if rng.random() < 0.3 else rng.randint(0, 9) for _ in range(n)]
    b = [rng.randint(0, 100) for _ in range(n)]
    c = [None if rng.random() < 0.3 else rng.randint(0, 9) for _ in range(n)]
    s = [None if rng.random() < 0.2 else rng.choice(["", "abc", "Banana", "xyz", "A"]) for _ in range(n)]
    return [a, b, c, s]


def _columns(vals):
    return [Column.from_values(v, t) for v, t in zip(vals, TYPES)]


def _expanded(cols, prog):
    """The program's input block as the executor builds it: physical columns + one `is None` companion per Option column."""
    out = list(cols)
    for j in sorted(prog.null_of):
        src = cols[prog.null_of[j]]
        pres = src.present()
        out.append(Column(T_BOOL, (~pres).astype(np.int64) if pres is not None else np.zeros(len(src), np.int64)))
    return out


def _compile(fn):
    sc = frontend.StageCompiler(TYPES, NAMES, OPTION)
    sc.add_map(fn, 100001)
    return sc.finish_memory()


def _oracle_rows(prog, res):
    """Python rows of an oracle result: visible columns with their companions folded back into None."""
    n_vis = len(prog.out_cols) - prog.hidden_out_cols
    cols = [res.values(k) for k in range(len(prog.out_cols))]
    for k, nof in enumerate(prog.out_null_of):
        if nof:
            cols[nof - 1] = [None if isnull else v for v, isnull in zip(cols[nof - 1], cols[k])]
    rows = list(zip(*cols[:n_vis]))
    return [r[0] if n_vis == 1 else r for r in rows]


@pytest.mark.parametrize("fn", UDFS, ids=[f.__name__ for f in UDFS])
def test_option_udfs_match_cpython(built, fn):
    n = 3000
    vals = _data(n, 7)
    cols = _columns(vals)
    prog = _compile(fn)
    assert prog.null_of == {4: 0, 5: 2, 6: 3}
    res = pyoracle.run_program(prog, _expanded(cols, prog), n)
    got = _oracle_rows(prog, res)
    exc = {int(e["row"]): int(e["code"]) for e in res.exceptions}
    j = 0
    for i in range(n):
        row = Row([v[i] for v in vals], NAMES)
        try:
            want = fn(row)
        except TypeError:
            assert exc.get(i) == C["TPLX_EC_TYPEERROR"], (fn.__name__, list(row.values if hasattr(row, "values") else []), exc.get(i))
            continue
        if i in exc:
            # the device may hand a row to the interpreter path (TypeError on a None it does not understand); CPython's answer stands
            assert exc[i] == C["TPLX_EC_TYPEERROR"]
            continue
        assert got[j] == want, (fn.__name__, [v[i] for v in vals], got[j], want)
        j += 1
    assert j == len(got)
    if fn.__name__ not in ("use_raises", "str_of_option", "less_than"):
        assert not exc, fn.__name__  # None-aware code paths: no row leaves the device


def test_descriptor_marks_companions(built):
    prog = _compile(UDFS[7])  # passthrough: Option columns flow to the output with their flags
    blob = prog.serialize()
    hdr = ir.HEADER_FMT
    import struct
    n_in = struct.unpack_from(hdr, blob, 0)[3]
    in_types = blob[struct.calcsize(hdr): struct.calcsize(hdr) + n_in]
    assert list(in_types) == TYPES + [C["TPLX_T_NULLOF"] | 0, C["TPLX_T_NULLOF"] | 2, C["TPLX_T_NULLOF"] | 3]
    assert prog.hidden_out_cols == 2 and prog.out_null_of == [0, 0, 0, 2, 3]
    st = backend.Stage(prog)  # descriptor validation (no device needed)
    st.close()
    bad = _compile(UDFS[7])
    bad.out_null_of = [0, 0, 0, 9, 3]
    with pytest.raises(backend.GpuBackendError):
        backend.Stage(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("jit", ["0", "2"])
def test_gpu_option_udfs_equal_oracle(gpu, jit, monkeypatch):
    """Physical columns + validity bitmaps in, validity bitmaps out; bit-equal to the oracle run over the expanded block."""
    monkeypatch.setenv("TPLX_JIT", jit)
    monkeypatch.setenv("TPLX_JIT_SYNC", "1")
    for n in (1, 777, 40000):
        vals = _data(n, n)
        cols = _columns(vals)
        for fn in UDFS:
            prog = _compile(fn)
            ora = pyoracle.run_program(prog, _expanded(cols, prog), n)
            st = backend.Stage(prog)
            res = st.run_host(0, cols, n)
            info = res.info
            assert int(info.n_out_rows) == ora.n_out and int(info.n_exceptions) == len(ora.exceptions), fn.__name__
            want = _oracle_rows(prog, ora)
            outs = [c.to_values() for c in res.columns()]
            got = list(zip(*outs)) if len(outs) > 1 else outs[0]
            assert list(got) == want, fn.__name__
            exc = res.exceptions()
            for f in ("row", "row_no", "code", "op_id"):
                assert np.array_equal(exc[f], ora.exceptions[f]), (fn.__name__, f)
            res.free()
            st.close()


@pytest.mark.gpu
def test_context_none_goldens(gpu):
    """The reference's Python-level goldens for None values (python/tests/test_nulls.py:23-45, test_is.py:24-40, README.md:29-34),
    now computed on the device: rows with None stay in the normal case as Option[T] columns."""
    import tuplex_b200 as tuplex
    c = tuplex.Context()
    ds = c.parallelize([1, None])
    assert ds.map(lambda x: x == None).collect() == [False, True]  # noqa: E711
    assert ds.map(lambda x: x != None).collect() == [True, False]  # noqa: E711
    assert c.parallelize([None, None]).map(lambda x: x == None).collect() == [True, True]  # noqa: E711
    assert c.parallelize([None, None]).map(lambda x: x != None).collect() == [False, False]  # noqa: E711
    assert c.parallelize([None, None]).map(lambda x: x == 42).collect() == [False, False]
    assert c.parallelize([None, None]).map(lambda x: x != 42).collect() == [True, True]
    assert c.parallelize([True, False, False, True]).map(lambda x: x is None).collect() == [False] * 4
    assert c.parallelize([None, None, None]).map(lambda x: x is not None).collect() == [False] * 3
    assert c.parallelize([None, True, False]).map(lambda x: x is not None).collect() == [False, True, True]
    before = c.metrics.exceptions
    d = c.parallelize([1, 2, None, 4]).map(lambda x: (x, x * x))
    assert d.collect() == [(1, 1), (2, 4), (4, 16)]          # README: the None row raises TypeError and is dropped
    assert c.metrics.exceptions == before + 1 and sum(d.exception_counts.values()) == 1
    # None-aware UDFs keep every row on the device, values and None flow through filters, withColumn and selectColumns
    rows = [(i, None if i % 3 == 0 else "s%d" % i, None if i % 5 == 0 else float(i)) for i in range(2000)]
    ds = c.parallelize(rows, columns=["k", "s", "f"])
    before = c.metrics.exceptions
    got = (ds.withColumn("n", lambda x: 0 if x["s"] is None else len(x["s"]))
             .filter(lambda x: x["f"] is not None)
             .selectColumns(["k", "s", "n", "f"]).collect())
    exp = [(k, s, 0 if s is None else len(s), f) for k, s, f in rows if f is not None]
    assert got == exp and c.metrics.exceptions == before
    # a resolver for the rows that do use None
    got = c.parallelize([3, None, 5]).map(lambda x: x + 1).resolve(TypeError, lambda x: -1).collect()
    assert got == [4, -1, 6]
    # the old behaviour (rows with None take the interpreter path) stays available and gives the same rows
    c2 = tuplex.Context({"tuplex.gpu.optionColumns": False})
    assert c2.parallelize([1, 2, None, 4]).map(lambda x: (x, x * x)).collect() == [(1, 1), (2, 4), (4, 16)]
    assert c2.parallelize([1, None]).map(lambda x: x == None).collect() == [False, True]  # noqa: E711


# ---- Option[T] fields in the reference's row format (K5, Serializer.cc:1016-1117) ---------------------------------------------
def _opt_cols(vals, types, option):
    cols = []
    for c, (v, t) in enumerate(zip(vals, types)):
        col = Column.from_values(v, t)
        if c in option and col.valid is None:  # statically Option[T] although this block holds no None
            col.valid = backend.pack_valid(np.ones(len(v), bool))
        cols.append(col)
    return cols


def test_oracle_row_bytes_with_option_fields_hand_derived(built):
    """(a: Option[i64] = None, s: str = 'ab', t: Option[str] = None, u: Option[str] = 'x', k: i64 = 7), bytes derived by hand from
    Serializer::serialize: bitmap word (bit k = k-th Option field is None) | 5 slots | var-len total | payload; a None string keeps the
    offset of the next var field with size 0 (appendWithoutInference(option<string>), Serializer.cc:313-338)."""
    import struct
    cols = _opt_cols([[None], ["ab"], [None], ["x"], [7]], [T_I64, T_STR, T_STR, T_STR, T_I64], {0, 2, 3})
    (part,) = pyoracle.to_partitions(cols, 1, 1 << 20)
    row = struct.pack("<q", 0b011)                    # a and t are None, u is present
    row += struct.pack("<q", 0)                       # a
    row += struct.pack("<q", 40 | (3 << 32))          # s: payload at slot address + 40, 'ab\\0'
    row += struct.pack("<q", 35)                      # t: None, size 0, where the next var field starts (51 - 16)
    row += struct.pack("<q", 27 | (2 << 32))          # u: 'x\\0' at 51 - 24
    row += struct.pack("<q", 7)                       # k
    row += struct.pack("<q", 5) + b"ab\0x\0"          # var-len total + payload
    assert part == struct.pack("<q", 1) + row and len(row) == 61


@pytest.mark.gpu
def test_gpu_row_format_with_option_fields(gpu):
    """K5 both ways and the exception partition with Option fields: byte-identical to the oracle."""
    n = 5000
    vals = _data(n, 99)
    cols = _opt_cols(vals, TYPES, set(OPTION))
    parts = pyoracle.to_partitions(cols, n, 64 << 10)
    assert len(parts) > 2
    # partitions -> column block (bitmaps -> validity) -> a pass-through stage -> columns + validity back
    blk = backend.Block.from_partitions(0, parts, TYPES, OPTION)
    sc = frontend.StageCompiler(TYPES, NAMES, OPTION)
    sc.add_map(lambda x: (x["a"], x["b"], x["c"], x["s"]), 100001)
    st = backend.Stage(sc.finish_memory())
    res = st.run(blk)
    got = [c.to_values() for c in res.columns()]
    assert got == vals
    # result -> partitions (validity -> bitmaps)
    assert res.partitions(64 << 10) == parts
    res.free()
    st.close()
    # exception rows keep their Option fields: the original input row with its bitmap
    sc = frontend.StageCompiler(TYPES, NAMES, OPTION)
    sc.add_map(lambda x: x["a"] + x["b"], 100001)
    prog = sc.finish_memory()
    st = backend.Stage(prog)
    res = st.run(blk)
    ora = pyoracle.run_program(prog, _expanded(cols, prog), n)
    exc = res.exceptions()
    assert len(exc) == sum(v is None for v in vals[0]) and np.array_equal(exc["row"], ora.exceptions["row"])
    assert res.exception_partition() == pyoracle.exception_partition(cols, ora.exceptions)
    res.free()
    st.close()
    blk.free()
    # a left join's nullable columns serialise with the bitmap as well
    left = [Column.from_values([1, 2, 3, 4], T_I64), Column.from_values(["a", "b", "c", "d"], T_STR)]
    right = [Column.from_values([2, 4], T_I64), Column.from_values(["two", "four"], T_STR), Column.from_values([20, 40], T_I64)]
    lb, rb = backend.Block.upload(0, left, 4), backend.Block.upload(0, right, 2)
    jn = backend.Join(rb, [T_I64, T_STR, T_I64], 0)
    jr = jn.probe(lb, [T_I64, T_STR], 0, left_outer=True)
    want = _opt_cols([["a", "b", "c", "d"], [1, 2, 3, 4], [None, "two", None, "four"], [None, 20, None, 40]], [T_STR, T_I64, T_STR, T_I64], {2, 3})
    assert jr.partitions(1 << 20) == pyoracle.to_partitions(want, 4, 1 << 20)
    jr.free()
    jn.free()
    lb.free()
    rb.free()
