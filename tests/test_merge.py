"""K9 (csrc/merge.cuh): resolved rows return to the slots of their exception records on the device (ResolveTask::executeInOrder,
tuplex/core/src/physical/ResolveTask.cc:878-1258) — against the host restatement `dataset._merge_by_rowno`, and through `tocsv`."""
import numpy as np
import pytest

from tuplex_b200 import backend, frontend
from tuplex_b200.backend import Column
from tuplex_b200.dataset import _merge_by_rowno
from tuplex_b200.ir import T_I64, T_STR


def _stage():
    sc = frontend.StageCompiler([T_I64, T_STR], ["k", "s"])
    sc.add_map(lambda x: (100 // x["k"], x["s"].upper() + "!", x["k"]), 100001)   # ZeroDivisionError for k == 0
    sc.add_filter(lambda a, b, c: c % 5 != 1, 100002)                             # rows leave, slots shift
    return sc.finish_memory()


def test_merge_closed_form_equals_reference_walk():
    """K9's closed form (normal row j -> j + #{m : a_m <= j}, resolved row m -> a_m + m, a_m = row number - exceptions before it) and the
    host restatement `_merge_by_rowno` against the oracle's literal walk of ResolveTask::executeInOrder (oracle/merge_oracle.c)."""
    from oracle import pyoracle
    rng = np.random.default_rng(2)
    for trial in range(200):
        n_norm, n_exc = int(rng.integers(0, 60)), int(rng.integers(0, 25))
        # a task's output stream: n_norm + n_exc slots, the exceptions sit at random distinct slots
        slots = np.sort(rng.choice(n_norm + n_exc, n_exc, replace=False)) if n_exc else np.zeros(0, np.int64)
        resolved = rng.random(n_exc) < 0.6
        want = pyoracle.merge_sources(n_norm, slots, resolved)
        a = np.array([int(slots[k]) - k for k in range(n_exc) if resolved[k]], np.int64)
        closed = np.empty(n_norm + len(a), np.int64)
        for j in range(n_norm):
            closed[j + int(np.searchsorted(a, j, side="right"))] = j
        for m, am in enumerate(a.tolist()):
            closed[am + m] = ~m
        assert closed.tolist() == want.tolist()
        exc = np.zeros(n_exc, dtype=backend.EXC_DTYPE)
        exc["row_no"] = slots
        res_rows = [(0, int(slots[k]), "R%d" % m) for m, k in enumerate([k for k in range(n_exc) if resolved[k]])]
        host = _merge_by_rowno(["n%d" % j for j in range(n_norm)], exc, res_rows)
        assert host == ["n%d" % v if v >= 0 else "R%d" % ~v for v in want.tolist()]


@pytest.mark.gpu
@pytest.mark.parametrize("first_row_no", [0, 1000])
def test_device_merge_equals_host_merge(gpu, first_row_no):
    rng = np.random.default_rng(4)
    for n in (1, 50, 20000):
        k = rng.integers(0, 7, n)
        s = ["w%d" % (v % 11) * int(v % 3) for v in rng.integers(0, 1000, n).tolist()]
        cols = [Column.from_values(k.tolist(), T_I64), Column.from_values(s, T_STR)]
        st = backend.Stage(_stage())
        res = st.run_host(0, cols, n, first_row_no)
        exc = res.exceptions()
        assert len(exc) == int((k == 0).sum())
        normal = list(zip(*[c.to_values() for c in res.columns()]))
        # resolver: every third exception stays unresolved, the others produce a row of the output schema
        resolved = [(int(e["row"]), int(e["row_no"]), (-1, "resolved-%d" % int(e["row"]), int(e["row"]))) for j, e in enumerate(exc) if j % 3 != 2]
        exc_local = exc.copy()
        exc_local["row_no"] -= first_row_no
        want = _merge_by_rowno(normal, exc_local, [(r, no - first_row_no, v) for r, no, v in resolved])
        rows = [v for _, _, v in resolved]
        blk = backend.Block.upload(0, [Column.from_values([r[0] for r in rows], T_I64), Column.from_values([r[1] for r in rows], T_STR),
                                       Column.from_values([r[2] for r in rows], T_I64)], len(rows))
        merged = res.merge_resolved(blk, [no for _, no, _ in resolved], first_row_no)
        assert int(merged.info.n_out_rows) == len(want) and int(merged.info.n_exceptions) == 0
        got = list(zip(*[c.to_values() for c in merged.columns()]))
        assert got == want
        # the merged result feeds the device writers like a stage result
        txt = merged.csv_bytes()
        assert txt.decode().splitlines() == ["%d,%s,%d" % r for r in want]
        parts = merged.partitions(1 << 20)
        assert sum(int(np.frombuffer(p[:8], "<i8")[0]) for p in parts) == len(want)
        if resolved:
            with pytest.raises(backend.GpuBackendError):
                res.merge_resolved(blk, [no + 1 for _, no, _ in resolved][::-1], first_row_no)  # not ascending / not exception slots
        merged.free()
        blk.free()
        res.free()
        st.close()


@pytest.mark.gpu
def test_tocsv_with_resolved_rows_stays_on_the_device(gpu, tmp_path, monkeypatch):
    import tuplex_b200 as tuplex
    calls = []
    orig = backend.Result.merge_resolved
    monkeypatch.setattr(backend.Result, "merge_resolved", lambda self, *a, **k: (calls.append(1), orig(self, *a, **k))[1])
    n = 30000
    rng = np.random.default_rng(8)
    data = [(int(a), "t%d" % i) for i, a in enumerate(rng.integers(0, 20, n))]
    c = tuplex.Context({"tuplex.gpu.blockRows": 8192})
    ds = (c.parallelize(data, columns=["a", "t"]).map(lambda x: (1000 // x["a"], x["t"]))
          .resolve(ZeroDivisionError, lambda x: (-1, x["t"])))
    out = tmp_path / "out.csv"
    ds.tocsv(str(out), header=False)
    want = ["%d,%s" % ((1000 // a) if a else -1, t) for a, t in data]
    assert out.read_text().splitlines() == want
    assert len(calls) == 4  # one merge per block, all on the device
    # without a resolver the slots stay empty and the rows are counted as exceptions
    ds2 = c.parallelize(data, columns=["a", "t"]).map(lambda x: (1000 // x["a"], x["t"]))
    out2 = tmp_path / "out2.csv"
    ds2.tocsv(str(out2), header=False)
    assert out2.read_text().splitlines() == [w for w, (a, _) in zip(want, data) if a]
    assert sum(ds2.exception_counts.values()) == sum(1 for a, _ in data if a == 0)
