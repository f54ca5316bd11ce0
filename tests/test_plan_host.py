"""Host-side plan logic that needs no device: column names from the logical plan (nothing is executed), and resolve() / ignore()
leaving the parent DataSet untouched (the reference adds a separate operator node, python/tuplex/dataset.py:344-389)."""
from tuplex_b200.backend import Column
from tuplex_b200.dataset import DataSet, Source
from tuplex_b200.ir import T_I64, T_STR


class _NoExecCtx:
    """A context that cannot execute anything: every attribute access fails."""
    def __getattr__(self, name):
        raise AssertionError(f"plan-only operation touched the context ({name}): it executed the pipeline")


def _ds():
    src = Source([Column.from_values([1, 2], T_I64), Column.from_values(["a", "b"], T_STR)], ["x", "s"], 2, None, [], 2)
    return DataSet(_NoExecCtx(), src)


def test_columns_do_not_execute():
    ds = _ds()
    assert ds.columns == ["x", "s"]
    assert ds.withColumn("y", lambda r: r["x"] + 1).columns == ["x", "s", "y"]
    assert ds.withColumn("y", lambda r: r["x"] + 1).selectColumns(["y", "s"]).renameColumn("s", "t").columns == ["y", "t"]
    assert ds.filter(lambda r: r["x"] > 1).mapColumn("s", lambda v: v.upper()).columns == ["x", "s"]
    assert ds.map(lambda r: {"a": r["x"], "b": r["s"]}).columns == ["a", "b"]
    assert ds.map(lambda r: (r["x"], r["s"], 3)).columns == [None, None, None]
    assert ds.aggregateByKey(lambda a, b: a + b, lambda a, r: a + r["x"], 0, ["s"]).columns == ["s", None]


def test_resolve_and_ignore_leave_the_parent_plan_alone():
    a = _ds().map(lambda r: r["x"] // (r["x"] - 1))
    b = a.resolve(ZeroDivisionError, lambda r: -1)
    c = a.ignore(ZeroDivisionError)
    assert a._ops[-1].resolvers == [] and a._ops[-1].ignores == []
    assert len(b._ops[-1].resolvers) == 1 and b._ops[-1].ignores == []
    assert c._ops[-1].resolvers == [] and c._ops[-1].ignores == [ZeroDivisionError]
    assert a._ops[-1].id == b._ops[-1].id == c._ops[-1].id  # same operator: exception counts are keyed by its id
    d = b.resolve(ValueError, lambda r: -2)
    assert len(b._ops[-1].resolvers) == 1 and len(d._ops[-1].resolvers) == 2


def test_types_and_text_source(tmp_path):
    """DataSet.types from the plan (python/tuplex/dataset.py:374-382) and Context.text (python/tuplex/context.py:367-387), no device."""
    import typing
    import tuplex_b200 as tuplex
    c = tuplex.Context()
    ds = c.parallelize([(1, "a", None), (2, "b", 3.5)], columns=["i", "s", "f"])
    assert ds.types == [int, str, typing.Optional[float]]
    assert ds.map(lambda x: (x["i"] * 2, x["s"].upper(), x["f"] is None)).types == [int, str, bool]
    assert ds.withColumn("g", lambda x: None if x["i"] > 1 else x["s"]).types == [int, str, typing.Optional[float], typing.Optional[str]]
    p = tmp_path / "t.txt"
    p.write_text("hello\nNULL\n\nworld\r\n")
    t = c.text(str(p), null_values=["NULL"])
    assert t._source.cols[0].to_values() == ["hello", None, "", "world"] and t.types == [typing.Optional[str]]
    assert c.text(str(p))._source.cols[0].to_values() == ["hello", "NULL", "", "world"]
