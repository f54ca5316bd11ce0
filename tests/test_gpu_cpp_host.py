"""The C++ host mirror (tuplex_b200/host/gpu_backend.{h,cc}: IBackend::execute over the C ABI) end to end:
reference-format partitions in, reference-format output and exception partitions out, byte-identical to the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from tuplex_b200 import frontend
from tuplex_b200.backend import Column
from tuplex_b200.ir import T_I64, T_STR
from oracle import pyoracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_gpu_backend_execute(gpu, tmp_path):
    rng = np.random.default_rng(8)
    n = 40_000
    a = rng.integers(-30, 30, n, dtype=np.int64)
    s = ["id%d,%d" % (v, v * v) for v in a]
    cols = [Column(T_I64, a), Column.from_values(s, T_STR)]
    sc = frontend.StageCompiler([T_I64, T_STR], ["a", "s"])
    sc.add_with_column("q", lambda x: 1000 // x['a'], 100001)          # ZeroDivisionError rows
    sc.add_with_column("t", lambda x: x['s'][x['s'].find(',') + 1:].upper() + '#', 100002)
    sc.add_filter(lambda x: x['q'] % 3 != 1, 100003)
    prog = sc.finish_memory()
    psize = 256 << 10
    in_parts = pyoracle.to_partitions(cols, n, psize)   # what TransformStage::inputPartitions() would hold
    assert len(in_parts) > 3
    desc = tmp_path / "stage.bin"
    desc.write_bytes(prog.serialize())
    part_files = []
    for i, p in enumerate(in_parts):
        f = tmp_path / f"in{i}.bin"
        f.write_bytes(p)
        part_files.append(str(f))
    exe = os.path.join(ROOT, "tuplex_b200", "lib", "tplx_host_run")
    prefix = str(tmp_path / "res")
    r = subprocess.run([exe, str(desc), "0,3", str(psize), prefix] + part_files, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    ora = pyoracle.run_program(prog, cols, n)
    assert info["out_rows"] == ora.n_out and info["exceptions"] == len(ora.exceptions) and info["exceptions"] > 0

    class OC:
        def __init__(self, t, d, o): self.type, self.data, self.offsets = t, d, o
    want_parts = pyoracle.to_partitions([OC(*c) for c in ora.columns], ora.n_out, psize)
    assert info["out_partitions"] == len(want_parts)
    for i, w in enumerate(want_parts):
        assert open(f"{prefix}.out{i}", "rb").read() == w, f"output partition {i}"
    assert open(prefix + ".exc", "rb").read() == pyoracle.exception_partition(cols, ora.exceptions)
    # stage-level failure surfaces as an error, not a crash
    bad = tmp_path / "bad.bin"
    bad.write_bytes(prog.serialize()[:-16])
    r2 = subprocess.run([exe, str(bad), "0,3", str(psize), prefix] + part_files, capture_output=True, text=True, timeout=120)
    assert r2.returncode == 1 and "stage descriptor" in r2.stderr


def _decode_partition(part: bytes, types):
    """rows of a reference-format partition (Partition.h:130-139 + Serializer.cc:1016-1117): int64 numRows, then per row one 8-byte slot
    per field (var-len: offset from the slot | size << 32), the 8-byte var-len total when the schema has strings, the payload."""
    import struct
    n = struct.unpack_from("<q", part, 0)[0]
    p = 8
    has_var = any(t == T_STR for t in types)
    rows = []
    for _ in range(n):
        vals = []
        for f, t in enumerate(types):
            slot = p + 8 * f
            raw = struct.unpack_from("<Q", part, slot)[0]
            if t == T_STR:
                off, size = raw & 0xFFFFFFFF, raw >> 32
                vals.append(part[slot + off: slot + off + size - 1].decode())
            elif t == 1:
                vals.append(struct.unpack_from("<d", part, slot)[0])
            else:
                vals.append(struct.unpack_from("<q", part, slot)[0])
        p += 8 * len(types)
        if has_var:
            p += 8 + struct.unpack_from("<q", part, p)[0]
        rows.append(tuple(vals))
    return rows


def _run_host(tmp_path, prog, cols, n, types, psize, devices=None):
    in_parts = pyoracle.to_partitions(cols, n, psize)
    desc = tmp_path / "stage.bin"
    desc.write_bytes(prog.serialize())
    files = []
    for i, p in enumerate(in_parts):
        f = tmp_path / f"in{i}.bin"
        f.write_bytes(p)
        files.append(str(f))
    exe = os.path.join(ROOT, "tuplex_b200", "lib", "tplx_host_run")
    prefix = str(tmp_path / ("res" + (devices or "").replace(",", "_")))
    cmd = [exe, str(desc), ",".join(str(t) for t in types), str(psize), prefix] + (["--devices", devices] if devices else []) + files
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1]), prefix, len(in_parts)


def test_cpp_host_tasks_aggregate_and_hash(gpu, tmp_path):
    """GpuBackend::execute beyond one memory task: several tasks (here on one GPU: --devices 0,0,0), exception row numbers stitched
    across tasks, the resolve hook, the aggregate endpoint (Q6 on the reference's lineitem fixture -> gtest golden) and the hash
    endpoint (AggregateTest.cc:249-364 goldens) — all through the C++ host over the C ABI."""
    import struct
    from tuplex_b200 import workloads, ir
    # ---- memory endpoint, 3 tasks: rows and exception records equal the single-task oracle run
    rng = np.random.default_rng(3)
    n = 30_000
    a = rng.integers(-20, 20, n, dtype=np.int64)
    s = ["id%d,%d" % (v, v * v) for v in a]
    cols = [Column(T_I64, a), Column.from_values(s, T_STR)]
    sc = frontend.StageCompiler([T_I64, T_STR], ["a", "s"])
    sc.add_with_column("q", lambda x: 1000 // x['a'], 100001)
    sc.add_filter(lambda x: x['q'] % 3 != 1, 100002)
    prog = sc.finish_memory()
    psize = 64 << 10
    info, prefix, n_in = _run_host(tmp_path, prog, cols, n, [T_I64, T_STR], psize, "0,0,0")
    ora = pyoracle.run_program(prog, cols, n)
    assert info["tasks"] == 3 and info["out_rows"] == ora.n_out and info["exceptions"] == len(ora.exceptions) > 0
    assert info["exceptions_handed_to_resolve"] == info["exceptions"]
    out_types = [T_I64, T_STR, T_I64]
    got_rows = []
    for i in range(info["out_partitions"]):
        got_rows += _decode_partition(open(f"{prefix}.out{i}", "rb").read(), out_types)
    want = list(zip(*[ora.values(c) for c in range(3)]))
    assert got_rows == want
    assert open(prefix + ".exc", "rb").read() == pyoracle.exception_partition(cols, ora.exceptions)
    # ---- aggregate endpoint: Q6 over the reference's SF0.01 fixture
    lcols = workloads.load_lineitem_fixture()
    ln = len(lcols[0].data)
    qp = workloads.q6_program()
    info, prefix, _ = _run_host(tmp_path, qp, lcols, ln, [0, 1, 1, 0], 256 << 10)
    (bits,) = struct.unpack("<q", open(prefix + ".agg", "rb").read())
    assert info["endpoint"] == 1 and bits & ((1 << 64) - 1) == pyoracle.run_program(qp, lcols, ln).acc_tree[0]
    assert abs(ir.bits_f64(bits) - 1193053.2252999984) <= 1e-4   # gtest golden (TPCH.cc:85-97)
    info2, prefix2, n_parts = _run_host(tmp_path, qp, lcols, ln, [0, 1, 1, 0], 256 << 10, "0,0")
    (bits2,) = struct.unpack("<q", open(prefix2 + ".agg", "rb").read())
    assert info2["tasks"] == 2 and abs(ir.bits_f64(bits2) - 1193053.2252999984) <= 1e-4
    # ---- hash endpoint: AggregateTest.cc goldens (scaled x2500), two tasks into one device table
    base = [("abc", 1), ("abc", -2), ("xyz", 4), ("abc", -2), ("xyz", 3), ("xyz", 3)]   # sums: abc -3, xyz 10
    rows = base * 2500
    kcols = [Column.from_values([r[0] for r in rows], T_STR), Column(T_I64, np.array([r[1] for r in rows], dtype=np.int64))]
    hs = frontend.StageCompiler([T_STR, T_I64], ["k", "v"])
    hp = hs.finish_hash(["k"], lambda a, x: a + x[1], lambda a, b: a + b, 0, 100001)
    for devs in (None, "0,0"):
        info, prefix, _ = _run_host(tmp_path, hp, kcols, len(rows), [T_STR, T_I64], 16 << 10, devs)
        got = []
        for i in range(info["hash_partitions"]):
            got += _decode_partition(open(f"{prefix}.hash{i}", "rb").read(), [T_STR, T_I64])
        assert sorted(got) == [("abc", -7500), ("xyz", 25000)] and info["out_rows"] == 2 and info["endpoint"] == 2


def test_cpp_host_hash_join(gpu, tmp_path):
    """GpuBackend::execute(GpuHashJoinStage&): reference-format partitions of both sides in, joined rows out as partitions (Option
    fields of a left join with their row bitmap), two tasks on one device; byte-identical to the oracle's rows serialised by the oracle."""
    from tuplex_b200 import backend
    rng = np.random.default_rng(12)
    n_probe, n_build = 30_000, 900
    pk = rng.integers(0, 1200, n_probe)
    probe = [Column(T_I64, pk.astype(np.int64)), Column.from_values(["p%d" % i for i in range(n_probe)], T_STR)]
    bk = rng.integers(0, 1200, n_build)
    build = [Column.from_values(["name-%d" % v for v in bk.tolist()], T_STR), Column(T_I64, bk.astype(np.int64)), Column(T_I64, np.arange(n_build, dtype=np.int64) * 10)]
    psize = 64 << 10
    pparts, bparts = pyoracle.to_partitions(probe, n_probe, psize), pyoracle.to_partitions(build, n_build, psize)
    files = []
    for tag, parts in (("b", bparts), ("p", pparts)):
        for i, p in enumerate(parts):
            f = tmp_path / f"{tag}{i}.bin"
            f.write_bytes(p)
            files.append(str(f))
    exe = os.path.join(ROOT, "tuplex_b200", "lib", "tplx_host_run")
    for left_outer in (False, True):
        op, ob = pyoracle.join_pairs(build[1], n_build, probe[0], n_probe, left_outer)
        pvals, bvals = [c.to_values() for c in probe], [c.to_values() for c in build]
        take = lambda vals, idx: [None if i < 0 else vals[i] for i in idx.tolist()]  # noqa: E731
        want_vals = [take(pvals[1], op), take(pvals[0], op), take(bvals[0], ob), take(bvals[2], ob)]   # | probe non-key | key | build non-key |
        want_cols = [Column.from_values(v, t) for v, t in zip(want_vals, [T_STR, T_I64, T_STR, T_I64])]
        if left_outer:
            for c in want_cols[2:]:  # Option[T] columns of a left join even where this run has no None
                if c.valid is None:
                    c.valid = backend.pack_valid(np.ones(len(op), bool))
        prefix = str(tmp_path / f"join{int(left_outer)}")
        cmd = [exe, "--join", "0,3", "0", "3,0,0", "1", str(1 if left_outer else 0), str(psize), prefix, "0,0", str(len(bparts))] + files
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        info = json.loads(r.stdout.strip().splitlines()[-1])
        assert info["out_rows"] == len(op) and info["tasks"] == 2
        got = b"".join(open(f"{prefix}.out{i}", "rb").read()[8:] for i in range(info["out_partitions"]))
        want = b"".join(p[8:] for p in pyoracle.to_partitions(want_cols, len(op), psize))
        assert got == want   # same rows, same bytes (partition boundaries differ: two tasks)
