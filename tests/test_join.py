"""Hash join (K8, csrc/join.cuh): the CPU oracle against the reference's own goldens (tuplex/test/core/JoinTest.cc), the CUDA build + probe
through the C ABI against the oracle (index pairs -> every output column, validity bitmaps), and the DataSet.join / leftJoin mirror."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tuplex_b200 import backend
from tuplex_b200.backend import Column
from tuplex_b200.ir import T_I64, T_STR, T_F64


# ---- oracle pinned to the reference's goldens (no GPU) ---------------------------------------------------------------------
def test_oracle_inner_join_str_golden():
    # JoinTest.cc:135-170 SimpleColumnBasedInnerJoinStr
    A = [(1, "one", 3), (1, "one", 4), (2, "two", 1), (4, "four", 2)]
    B = [(1, "one"), (10, "one"), (2, "two"), (3, "three")]
    assert po.join_rows(A, 1, B, 1, T_STR) == [(1, 3, "one", 1), (1, 3, "one", 10), (1, 4, "one", 1), (1, 4, "one", 10), (2, 1, "two", 2)]


def test_oracle_single_column_join_golden():
    # JoinTest.cc:230-246 InnerJoinSingleCol
    A = [("a",), ("b",), ("c",), ("d",), ("e",)]
    B = [("b",), ("d",)]
    assert po.join_rows(A, 0, B, 0, T_STR) == [("b",), ("d",)]


def test_oracle_int_joins_golden():
    # JoinTest.cc:416-460 SimpleIntJoins
    ds1 = [(1, 2, 3), (1, 2, 4), (2, 1, 1)]
    ds2 = [(1, "one"), (2, "two"), (3, "three")]
    assert po.join_rows(ds1, 1, ds2, 0, T_I64) == [(1, 3, 2, "two"), (1, 4, 2, "two"), (2, 1, 1, "one")]
    ds3 = [(1, 2, 4), (1, 2, 5), (2, 1, 8), (1, 10, 2)]
    assert po.join_rows(ds3, 1, ds2, 0, T_I64, left_outer=True) == [(1, 4, 2, "two"), (1, 5, 2, "two"), (2, 8, 1, "one"), (1, 2, 10, None)]
    assert po.join_rows(ds3, 1, ds2, 0, T_I64) == [(1, 4, 2, "two"), (1, 5, 2, "two"), (2, 8, 1, "one")]


def test_oracle_left_join_goldens():
    # JoinTest.cc:284-303 SimpleLeftJoin (selectColumns A, B, C afterwards), :305-325, :327-348
    assert po.join_rows([("abc", 20), ("def", 30)], 0, [(30, "abc"), (40, "xyz")], 1, T_STR, left_outer=True) == [(20, "abc", 30), (30, "def", None)]
    ds = [("ATL", "FRA", 20), ("FRA", "BOS", 10)]
    assert po.join_rows(ds, 0, [("ATL", "Atlanta")], 0, T_STR, left_outer=True) == [("FRA", 20, "ATL", "Atlanta"), ("BOS", 10, "FRA", None)]


def test_oracle_null_bucket_goldens():
    # JoinTest.cc:21-76 InnerJoinNullBucket, :77-133 InnerJoinInt64Option: None matches None
    dsA = [("abc", 42), (None, 84), ("xyz", 100)]
    dsB = [(None, -1), (None, -2)]
    assert po.join_rows(dsA, 0, dsB, 0, T_STR) == [(84, None, -1), (84, None, -2)]
    dsD = [(None, 84, "hello")]
    assert po.join_rows(dsD, 0, dsB, 0, T_STR) == [(84, "hello", None, -1), (84, "hello", None, -2)]
    iA = [(1, "abc"), (None, "def"), (2, "ghi")]
    assert po.join_rows(iA, 0, dsB, 0, T_I64) == [("def", None, -1), ("def", None, -2)]


def test_oracle_build_left_order():
    # JoinOperator::buildRight() false (left side smaller, inner join): the probe runs over the RIGHT rows, so they order the output
    # (PipelineBuilder.cc:2179-2198); JoinTest.cc:248-282 InnerJoinTwoTimes first join (4 flights vs 5 airports)
    ds = [("ATL", "FRA", 20), ("FRA", "BOS", 10), ("JFK", "ATL", 5), ("BOS", "JFK", 0)]
    ap = [("ATL", "Atlanta"), ("FRA", "Frankfurt"), ("JFK", "New York"), ("LAX", "Los Angeles"), ("TXL", "Berlin")]
    got = po.join_rows(ds, 0, ap, 0, T_STR, build_right=False)
    assert got == [("FRA", 20, "ATL", "Atlanta"), ("BOS", 10, "FRA", "Frankfurt"), ("ATL", 5, "JFK", "New York")]


# ---- CUDA build + probe against the oracle ----------------------------------------------------------------------------------
def _expected_columns(left, lk, right, rk, op, ob, build_first):
    """Assemble the expected output columns from the oracle's index pairs with numpy / python lists."""
    def gather(col: Column, idx, nullable):
        vals = col.to_values()
        return [None if i < 0 else vals[i] for i in idx.tolist()], nullable or col.valid is not None

    probe, build, pk, bk = (right, left, rk, lk) if build_first else (left, right, lk, rk)
    first, fk, fidx = (build, bk, ob) if build_first else (probe, pk, op)
    second, sk, sidx = (probe, pk, op) if build_first else (build, bk, ob)
    left_outer = bool((ob < 0).any())
    cols = [gather(c, fidx, False) for i, c in enumerate(first) if i != fk]
    cols.append(gather(probe[pk], op, False))
    cols += [gather(c, sidx, left_outer and not build_first) for i, c in enumerate(second) if i != sk]
    return cols


def _run_gpu_join(left, lk, right, rk, left_outer=False, build_first=False):
    probe, build, pk, bk = (right, left, rk, lk) if build_first else (left, right, lk, rk)
    bb = backend.Block.upload(0, build, len(build[0]))
    pb = backend.Block.upload(0, probe, len(probe[0]))
    j = backend.Join(bb, [c.type for c in build], bk)
    res = j.probe(pb, [c.type for c in probe], pk, left_outer=left_outer, build_first=build_first)
    cols = res.columns()
    info = res.info
    res.free()
    j.free()
    pb.free()
    bb.free()
    return cols, info


def _check(left, lk, right, rk, left_outer=False, build_first=False):
    probe, build, pk, bk = (right, left, rk, lk) if build_first else (left, right, lk, rk)
    op, ob = po.join_pairs(build[bk], len(build[0]), probe[pk], len(probe[0]), left_outer)
    exp = _expected_columns(left, lk, right, rk, op, ob, build_first)
    got, info = _run_gpu_join(left, lk, right, rk, left_outer, build_first)
    assert int(info.n_out_rows) == len(op)
    assert len(got) == len(exp)
    for k, (g, (ev, _)) in enumerate(zip(got, exp)):
        assert len(g) == len(op), k
        assert g.to_values() == ev, f"output column {k}"
    return len(op)


def _rand_cols(rng, n, n_keys, key_type, null_frac=0.0, with_str=True):
    ids = rng.integers(0, n_keys, n)
    if key_type == T_STR:
        keys = [f"k{v:05d}" if v % 7 else f"key-{v}-long-{'x' * int(v % 23)}" for v in ids.tolist()]
    else:
        keys = (ids * 1000003 - 17).tolist()
    if null_frac:
        keys = [None if rng.random() < null_frac else k for k in keys]
    cols = [Column.from_values(keys, key_type), Column.from_values(rng.integers(-10**9, 10**9, n).tolist(), T_I64)]
    if with_str:
        cols.append(Column.from_values([("s%d" % v) * int(v % 5) for v in rng.integers(0, 10**6, n).tolist()], T_STR))
    cols.append(Column.from_values(rng.random(n).tolist(), T_F64))
    return cols


@pytest.mark.gpu
def test_gpu_join_goldens(gpu):
    A = [Column.from_values([1, 1, 2, 4], T_I64), Column.from_values(["one", "one", "two", "four"], T_STR), Column.from_values([3, 4, 1, 2], T_I64)]
    B = [Column.from_values([1, 10, 2, 3], T_I64), Column.from_values(["one", "one", "two", "three"], T_STR)]
    got, _ = _run_gpu_join(A, 1, B, 1)
    rows = list(zip(*[c.to_values() for c in got]))
    assert rows == [(1, 3, "one", 1), (1, 3, "one", 10), (1, 4, "one", 1), (1, 4, "one", 10), (2, 1, "two", 2)]
    ds3 = [Column.from_values([1, 1, 2, 1], T_I64), Column.from_values([2, 2, 1, 10], T_I64), Column.from_values([4, 5, 8, 2], T_I64)]
    ds2 = [Column.from_values([1, 2, 3], T_I64), Column.from_values(["one", "two", "three"], T_STR)]
    got, _ = _run_gpu_join(ds3, 1, ds2, 0, left_outer=True)
    rows = list(zip(*[c.to_values() for c in got]))
    assert rows == [(1, 4, 2, "two"), (1, 5, 2, "two"), (2, 8, 1, "one"), (1, 2, 10, None)]
    dsA = [Column.from_values(["abc", None, "xyz"], T_STR), Column.from_values([42, 84, 100], T_I64)]
    dsB = [Column.from_values([None, None], T_STR), Column.from_values([-1, -2], T_I64)]
    got, _ = _run_gpu_join(dsA, 0, dsB, 0)
    assert list(zip(*[c.to_values() for c in got])) == [(84, None, -1), (84, None, -2)]


@pytest.mark.gpu
@pytest.mark.parametrize("key_type", [T_I64, T_STR])
@pytest.mark.parametrize("left_outer", [False, True])
def test_gpu_join_random_vs_oracle(gpu, key_type, left_outer):
    rng = np.random.default_rng(11 + key_type + 2 * left_outer)
    for n_probe, n_build, n_keys, nulls in [(1, 1, 1, 0.0), (1000, 37, 50, 0.0), (20011, 3001, 1500, 0.05), (50000, 20000, 40000, 0.0), (333, 0, 5, 0.0)]:
        left = _rand_cols(rng, n_probe, n_keys, key_type, nulls)
        right = _rand_cols(rng, n_build, n_keys, key_type, nulls, with_str=(n_build % 2 == 1))
        _check(left, 0, right, 0, left_outer)


@pytest.mark.gpu
def test_gpu_join_heavy_buckets_keep_build_order(gpu):
    """Buckets of 33 .. 5000 rows: the CTA-per-group ordering kernel and the warp-wide emit loop; output order = build-row order."""
    rng = np.random.default_rng(5)
    n_build = 12000
    ids = np.concatenate([np.full(5000, 1), np.full(2100, 2), np.full(40, 3), np.full(33, 4), rng.integers(10, 2000, n_build - 7173)])
    rng.shuffle(ids)
    right = [Column.from_values(ids.tolist(), T_I64), Column.from_values(np.arange(n_build).tolist(), T_I64),
             Column.from_values(["r%d" % i for i in range(n_build)], T_STR)]
    left = [Column.from_values([7, 1, 3, 99999, 4, 2, 1, 11], T_I64), Column.from_values(list("abcdefgh"), T_STR)]
    n = _check(left, 0, right, 0, left_outer=True)
    assert n >= 5000 * 2 + 2100 + 40 + 33
    got, _ = _run_gpu_join(left, 0, right, 0)
    pos = np.asarray(got[2].to_values())  # build row numbers: ascending inside every probe row's run
    keys = np.asarray(got[1].to_values())
    for k in (1, 2, 3, 4):
        run = pos[keys == k]
        per_probe = run.reshape(-1, int((ids == k).sum()))
        assert (np.diff(per_probe, axis=1) > 0).all()


@pytest.mark.gpu
def test_gpu_join_build_first(gpu):
    """Build side = LEFT dataset (JoinOperator::buildRight() false): probe rows (right) order the output, left columns come first."""
    rng = np.random.default_rng(9)
    left = _rand_cols(rng, 500, 300, T_STR)
    right = _rand_cols(rng, 4000, 300, T_STR)
    _check(left, 0, right, 0, build_first=True)
    ds = [Column.from_values(["ATL", "FRA", "JFK", "BOS"], T_STR), Column.from_values(["FRA", "BOS", "ATL", "JFK"], T_STR), Column.from_values([20, 10, 5, 0], T_I64)]
    ap = [Column.from_values(["ATL", "FRA", "JFK", "LAX", "TXL"], T_STR), Column.from_values(["Atlanta", "Frankfurt", "New York", "Los Angeles", "Berlin"], T_STR)]
    got, _ = _run_gpu_join(ds, 0, ap, 0, build_first=True)
    assert list(zip(*[c.to_values() for c in got])) == [("FRA", 20, "ATL", "Atlanta"), ("BOS", 10, "FRA", "Frankfurt"), ("ATL", 5, "JFK", "New York")]


@pytest.mark.gpu
def test_gpu_join_rejects_bad_arguments(gpu):
    a = [Column.from_values([1.5, 2.5], T_F64)]
    bb = backend.Block.upload(0, a, 2)
    with pytest.raises(backend.GpuBackendError):
        backend.Join(bb, [T_F64], 0)  # f64 keys are not hashable on this path
    b = [Column.from_values([1, 2], T_I64)]
    s = [Column.from_values(["x", "y"], T_STR)]
    ib, sb = backend.Block.upload(0, b, 2), backend.Block.upload(0, s, 2)
    j = backend.Join(ib, [T_I64], 0)
    with pytest.raises(backend.GpuBackendError):
        j.probe(sb, [T_STR], 0)  # key types differ
    with pytest.raises(backend.GpuBackendError):
        j.probe(ib, [T_I64], 0, left_outer=True, build_first=True)  # that would be a right join
    j.free()


# ---- DataSet.join / leftJoin (python/tuplex/dataset.py:384-498) ---------------------------------------------------------------
def test_join_plan_names_and_build_side():
    """Column names and the build side come from the logical plan alone (JoinOperator.cc:163-184, JoinOperator.h:62-69)."""
    import tuplex_b200 as tuplex
    c = tuplex.Context()
    ds = c.parallelize([("ATL", "FRA", 20), ("FRA", "BOS", 10), ("JFK", "ATL", 5), ("BOS", "JFK", 0)], columns=["Origin", "Dest", "Delay"])
    ap = c.parallelize([("ATL", "Atlanta"), ("FRA", "Frankfurt"), ("JFK", "New York"), ("LAX", "Los Angeles"), ("TXL", "Berlin")], columns=["Code", "Name"])
    small = c.parallelize([("ATL", "Atlanta"), ("FRA", "Frankfurt"), ("JFK", "New York")], columns=["Code", "Name"])
    j = ds.join(ap, "Origin", "Code", prefixes=(None, "Origin"))
    assert j.columns == ["Dest", "Delay", "Origin", "OriginName"]
    assert j._join.build_right() is False        # 4 rows vs 5: the left side is the smaller one
    assert ds.join(small, "Origin", "Code")._join.build_right() is True
    assert ds.leftJoin(ap, "Origin", "Code")._join.build_right() is True  # a left join always builds on the right
    j2 = j.join(ap, "Dest", "Code", prefixes=(None, "Dest"))
    assert j2.columns == ["Delay", "Origin", "OriginName", "Dest", "DestName"]
    assert j2._join.build_right() is True        # cost of a join = sum of its parents' (LogicalOperator.h:197-204): 9 >= 5
    assert j2.selectColumns(["Origin", "OriginName", "Dest", "DestName", "Delay"]).columns == ["Origin", "OriginName", "Dest", "DestName", "Delay"]
    with pytest.raises(ValueError):
        ds.join(ap, "nope", "Code")
    s = ds.join(ap, "Origin", "Code", suffixes=("_l", "_r"))
    assert s.columns == ["Dest_l", "Delay_l", "Origin_l", "Name_r"]


def test_interpreter_path_join_matches_oracle():
    """The host twin used for fallback rows follows the same order rules as the oracle."""
    from tuplex_b200 import dataset
    rng = np.random.default_rng(3)
    L = [(int(a), "l%d" % i) for i, a in enumerate(rng.integers(0, 20, 200))]
    R = [(int(a), "r%d" % i, float(i)) for i, a in enumerate(rng.integers(0, 25, 150))]
    for lo in (False, True):
        assert dataset._py_join_pairs(L, 0, R, 0, lo, True, 3) == po.join_rows(L, 0, R, 0, T_I64, left_outer=lo)
    assert dataset._py_join_pairs(L, 0, R, 0, False, False, 3) == po.join_rows(L, 0, R, 0, T_I64, build_right=False)


@pytest.mark.gpu
def test_context_join_goldens(gpu):
    import tuplex_b200 as tuplex
    c = tuplex.Context()
    # JoinTest.cc:135-191 SimpleColumnBasedInnerJoinStr, with a filter before / after the join
    dsA = c.parallelize([(1, "one", 3), (1, "one", 4), (2, "two", 1), (4, "four", 2)], columns=["a", "b", "c"])
    dsB = c.parallelize([(1, "one"), (10, "one"), (2, "two"), (3, "three")], columns=["x", "y"])
    assert dsA.join(dsB, "b", "y").collect() == [(1, 3, "one", 1), (1, 3, "one", 10), (1, 4, "one", 1), (1, 4, "one", 10), (2, 1, "two", 2)]
    assert dsA.filter(lambda a, b, c: a % 2 == 0).join(dsB, "b", "y").collect() == [(2, 1, "two", 2)]
    assert dsA.join(dsB, "b", "y").filter(lambda a, b, c, d: d > 5).collect() == [(1, 3, "one", 10), (1, 4, "one", 10)]
    # JoinTest.cc:230-246 InnerJoinSingleCol
    assert c.parallelize(["a", "b", "c", "d", "e"], columns=["colA"]).join(c.parallelize(["b", "d"], columns=["colA"]), "colA", "colA").collect() == ["b", "d"]
    # JoinTest.cc:248-282 InnerJoinTwoTimes (first join builds on the LEFT side: 4 flights vs 5 airports)
    ds = c.parallelize([("ATL", "FRA", 20), ("FRA", "BOS", 10), ("JFK", "ATL", 5), ("BOS", "JFK", 0)], columns=["Origin", "Dest", "Delay"])
    ap = c.parallelize([("ATL", "Atlanta"), ("FRA", "Frankfurt"), ("JFK", "New York"), ("LAX", "Los Angeles"), ("TXL", "Berlin")], columns=["Code", "Name"])
    small = c.parallelize([("ATL", "Atlanta"), ("FRA", "Frankfurt"), ("JFK", "New York")], columns=["Code", "Name"])
    for airports in (ap, small):
        res = (ds.join(airports, "Origin", "Code", prefixes=(None, "Origin")).join(airports, "Dest", "Code", prefixes=(None, "Dest"))
               .selectColumns(["Origin", "OriginName", "Dest", "DestName", "Delay"]).collect())
        assert res == [("ATL", "Atlanta", "FRA", "Frankfurt", 20), ("JFK", "New York", "ATL", "Atlanta", 5)]
    # JoinTest.cc:284-303 SimpleLeftJoin
    A = c.parallelize([("abc", 20), ("def", 30)], columns=["A", "B"])
    B = c.parallelize([(30, "abc"), (40, "xyz")], columns=["C", "D"])
    assert A.leftJoin(B, "A", "D").selectColumns(["A", "B", "C"]).collect() == [("abc", 20, 30), ("def", 30, None)]
    # JoinTest.cc:350-380 LeftJoinTwoTimes
    ds2 = c.parallelize([("ATL", "FRA", 20), ("FRA", "BOS", 10), ("JFK", "ATL", 5), ("BOS", "JFK", 12)], columns=["Origin", "Dest", "Delay"])
    res = (ds2.leftJoin(ap, "Origin", "Code", prefixes=(None, "Origin")).leftJoin(ap, "Dest", "Code", prefixes=(None, "Dest"))
           .selectColumns(["Origin", "OriginName", "Dest", "DestName", "Delay"]).collect())
    assert res == [("ATL", "Atlanta", "FRA", "Frankfurt", 20), ("FRA", "Frankfurt", "BOS", None, 10), ("JFK", "New York", "ATL", "Atlanta", 5),
                   ("BOS", None, "JFK", "New York", 12)]
    # JoinTest.cc:416-460 SimpleIntJoins
    ds1 = c.parallelize([(1, 2, 3), (1, 2, 4), (2, 1, 1)], columns=["a", "b", "c"])
    dsn = c.parallelize([(1, "one"), (2, "two"), (3, "three")], columns=["x", "y"])
    assert ds1.join(dsn, "b", "x").filter(lambda x: x[0] < 10).collect() == [(1, 3, 2, "two"), (1, 4, 2, "two"), (2, 1, 1, "one")]
    assert ds1.collect() == [(1, 2, 3), (1, 2, 4), (2, 1, 1)]
    ds3 = c.parallelize([(1, 2, 4), (1, 2, 5), (2, 1, 8), (1, 10, 2)], columns=["a", "b", "c"])
    assert ds3.leftJoin(dsn, "b", "x").collect() == [(1, 4, 2, "two"), (1, 5, 2, "two"), (2, 8, 1, "one"), (1, 2, 10, None)]
    # JoinTest.cc:21-133 null bucket: None matches None
    oa = c.parallelize([("abc", 42), (None, 84), ("xyz", 100)], columns=["a", "b"])
    ob = c.parallelize([(None, -1), (None, -2)], columns=["x", "y"])
    assert oa.join(ob, "a", "x").collect() == [(84, None, -1), (84, None, -2)]
    od = c.parallelize([(None, 84, "hello")], columns=["a", "b", "c"])
    assert od.join(ob, "a", "x").collect() == [(84, "hello", None, -1), (84, "hello", None, -2)]
    ia = c.parallelize([(1, "abc"), (None, "def"), (2, "ghi")], columns=["a", "b"])
    assert ia.join(ob, "a", "x").collect() == [("def", None, -1), ("def", None, -2)]


@pytest.mark.gpu
def test_context_join_with_fallback_rows_and_blocks(gpu):
    """Rows outside the normal case on either side join on the interpreter path and merge in order; several probe blocks."""
    import tuplex_b200 as tuplex
    from tuplex_b200 import dataset
    rng = np.random.default_rng(21)
    L = [(int(k), "l%d" % i) for i, k in enumerate(rng.integers(0, 300, 5000))]
    R = [(int(k), float(i)) for i, k in enumerate(rng.integers(0, 400, 700))]
    L[17] = ("seventeen", "odd")       # key of another type
    L[4000] = (5, 123)                 # payload of another type
    R[3] = (5, "five")                 # payload of another type on the build side
    R[650] = (1 << 70, 1.0)            # key beyond 64 bits
    c = tuplex.Context({"tuplex.gpu.blockRows": 1024})
    dl, dr = c.parallelize(L, columns=["k", "v"]), c.parallelize(R, columns=["k", "w"])
    for lo in (False, True):
        exp = dataset._py_join_pairs(L, 0, R, 0, lo, True, 2)
        got = (dl.leftJoin(dr, "k", "k") if lo else dl.join(dr, "k", "k")).collect()
        assert got == exp
    # inner join that builds on the left side (probe = right rows)
    small = c.parallelize(L[:300], columns=["k", "v"])
    assert small.join(dr, "k", "k").collect() == dataset._py_join_pairs(L[:300], 0, R, 0, False, False, 2)


@pytest.mark.gpu
def test_context_join_sharded_over_tasks(gpu):
    """tuplex.gpu.devices lists the device twice: two tasks, each builds its own table and probes its contiguous run of blocks
    (broadcast join, no exchange); the concatenation in task order equals the single-task result."""
    import tuplex_b200 as tuplex
    rng = np.random.default_rng(33)
    L = [(int(k), "l%d" % i, float(i)) for i, k in enumerate(rng.integers(0, 500, 9000))]
    R = [(int(k), "r%d" % i) for i, k in enumerate(rng.integers(0, 600, 1200))]
    one = tuplex.Context({"tuplex.gpu.blockRows": 1000})
    two = tuplex.Context({"tuplex.gpu.blockRows": 1000, "tuplex.gpu.devices": "0,0"})
    for ctx_ in (one, two):
        dl, dr = ctx_.parallelize(L, columns=["k", "v", "f"]), ctx_.parallelize(R, columns=["k", "w"])
        assert dl.leftJoin(dr, "k", "k").collect() == po.join_rows(L, 0, R, 0, T_I64, left_outer=True)
        assert dl.join(dr, "k", "k", prefixes=("l_", "r_")).columns == ["l_v", "l_f", "l_k", "r_w"]
