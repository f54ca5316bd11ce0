"""GPU parity for the CSV source (K6), through the C ABI: tplx_gpu_csv_upload / tplx_gpu_csv_parse.

The parsed block is read back through an identity stage (tplx_gpu_stage_run) and compared bit for bit with the CPU
oracle (oracle/csv_oracle.c, pinned in tests/test_csv_oracle.py): values, string bytes and offsets, the block-row ->
data-row map, and the list of rows handed to the interpreter path (row, code, byte range).
"""
import gzip
import hashlib
import os
import random

import numpy as np
import pytest

from tuplex_b200 import backend, frontend, ir, workloads
from tuplex_b200.ir import T_BOOL, T_F64, T_I64, T_STR
from oracle import pyoracle as po
from csv_helpers import T_SKIP, Parsed, assert_same_parse, gen_csv

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
I, F, B, S, X = T_I64, T_F64, T_BOOL, T_STR, T_SKIP


def gpu_parse(data: bytes, types, **kw) -> Parsed:
    buf = backend.CsvBuffer(0, data)
    p = buf.parse(types, **kw)
    info = p.info
    out_types = [t for t in types if t != X]
    cols = []
    if out_types:
        sc = frontend.StageCompiler(out_types, [None] * len(out_types))
        sc.add_map(lambda x: x, 100001)
        st = backend.Stage(sc.finish_memory())
        res = st.run(p.block)
        assert int(res.info.n_out_rows) == int(info.n_normal) and int(res.info.n_exceptions) == 0
        for c, t in enumerate(out_types):
            col = res.column(c)
            cols.append((col.data.tobytes(), col.offsets) if t == T_STR else (col.data.view(np.float64) if t == T_F64 else col.data))
        res.free()
        st.close()
    bad = [(int(b["row"]), int(b["code"]), int(b["line_start"]), int(b["line_end"])) for b in p.bad_rows()]
    out = Parsed(int(info.n_rows), cols, out_types, p.rowmap(), bad, int(info.sequential_rows))
    assert int(info.n_normal) + int(info.n_bad) == int(info.n_rows)
    p.free()
    buf.free()
    return out


def test_reference_vectors_on_gpu(gpu):
    from test_csv_oracle import VECTORS, _values
    for text, types, ok, expect in VECTORS:
        r = gpu_parse(text.encode(), types, null_values=[])
        assert r.n_rows == 1
        if ok:
            assert not r.bad and _values(r) == expect, text
        else:
            assert len(r.bad) == 1 and r.bad[0][1] == 70, text


def test_fuzz_equals_oracle(gpu):
    rng = random.Random(23)
    seq = 0
    for it in range(120):
        ncols = rng.randint(1, 7)
        types = [rng.choice([I, F, B, S, S, X]) for _ in range(ncols)]
        data = gen_csv(rng, rng.choice([0, 1, 7, 300, 3000]), types, dirty=rng.choice([0.0, 0.05, 0.3]),
                       weird_quotes=rng.choice([0.0, 0.0, 0.01]))
        kw = dict(header=rng.random() < 0.5, null_values=rng.choice([[], [""], ["", "NULL"]]))
        a = gpu_parse(data, types, **kw)
        assert_same_parse(a, po.csv_parse(data, types, **kw), what=(it, data[:120]))
        seq += a.sequential
    assert seq > 0  # irregular quoting went through csv_rows_sequential on the device


def test_edge_inputs(gpu):
    for data in (b"", b"\n", b"\r\n\r\n", b"a", b"a,b", b'"', b'""', b'"a\nb"', b",", b",\n,", b"x" * 70000, b'"' + b"y" * 40000 + b'"\n1',
                 (b"1,2\n" * 5000) + b'3,"4', b"a\r\nb\rc\nd"):
        for types in ([S], [S, S], [I, S]):
            for header in (False, True):
                a = gpu_parse(data, types, header=header)
                assert_same_parse(a, po.csv_parse(data, types, header=header), what=(data[:40], types, header))


def test_tile_boundaries(gpu):
    """quotes, newlines and \\r\\n pairs placed across 64-byte span and 16 KB tile boundaries"""
    rng = random.Random(5)
    for it in range(30):
        pad = rng.choice([60, 62, 63, 64, 16380, 16383, 16384, 32767])
        body = rng.choice([b'"q,\n"', b"\r\n", b'""', b'a,"b\r\nc"\r\n', b"\n\n\n"])
        data = b"x" * pad + body + b"tail,1\n" + b"u,v\n" * rng.randint(0, 5000) + b'"last ""one""",2'
        a = gpu_parse(data, [S, S])
        assert_same_parse(a, po.csv_parse(data, [S, S]), what=(pad, body))


def _zillow_csv():
    return gzip.open(os.path.join(HERE, "golden", "zillow_noexc.csv.gz"), "rb").read()


ZILLOW_FILE_TYPES = [S, S, S, S, F, S, S, X, S, X]  # title..url with projection pushdown; postal_code inferred f64


def test_zillow_csv_to_golden_md5(gpu):
    """reference fixture CSV -> device parse -> Z1 stage -> the md5 of the reference's own C++ / Python baselines"""
    data = _zillow_csv()
    buf = backend.CsvBuffer(0, data)
    p = buf.parse(ZILLOW_FILE_TYPES, header=True)
    assert int(p.info.n_rows) == 32661 and int(p.info.n_bad) == 0 and int(p.info.sequential_rows) == 0
    st = backend.Stage(workloads.zillow_program())
    res = st.run(p.block)
    vals = [c.to_values() for c in res.columns()]
    txt = workloads.rows_to_csv(vals, workloads.ZILLOW_OUT)
    assert hashlib.md5(txt).hexdigest() == "4d5ca0263b1a5058341a369116dee83a"


def test_zillow_csv_replicated(gpu):
    """full-size property: k copies of the fixture body parse to k * 32,661 rows and k * 577 output rows"""
    data = _zillow_csv()
    head, body = data.split(b"\n", 1)
    k = 40  # ~270 MB of CSV
    big = head + b"\n" + body * k
    buf = backend.CsvBuffer(0, big)
    p = buf.parse(ZILLOW_FILE_TYPES, header=True)
    assert int(p.info.n_rows) == 32661 * k and int(p.info.n_bad) == 0
    st = backend.Stage(workloads.zillow_program())
    res = st.run(p.block)
    assert int(res.info.n_out_rows) == 577 * k and int(res.info.n_exceptions) == 0
    vals = [c.to_values() for c in res.columns()]
    golden = workloads.zillow_golden_csv().decode().split("\n")[1:-1]
    for j in (0, 17, k - 1):
        assert workloads.rows_to_csv([v[j * 577:(j + 1) * 577] for v in vals], None).decode().split("\n")[:-1] == golden


# ---- CSV sink (K7) through the C ABI: tplx_gpu_result_csv ---------------------------------------------------------------
def test_sink_zillow_output_file_bytes(gpu):
    cols, n = workloads.load_zillow_fixture()
    st = backend.Stage(workloads.zillow_program())
    res = st.run_host(0, cols, n)
    body = workloads.zillow_golden_csv().split(b"\n", 1)[1]
    assert res.csv_bytes() == body  # byte-identical to the reference baselines' output rows
    assert hashlib.md5(",".join(workloads.ZILLOW_OUT).encode() + b"\n" + res.csv_bytes()).hexdigest() == "4d5ca0263b1a5058341a369116dee83a"


def test_sink_fuzz_equals_oracle(gpu):
    from tuplex_b200.backend import Column
    rng = random.Random(77)
    words = ["", "a", "x,y", 'q"q', "l\nb", "r\rb", "plain text", "é", '""', ";", "tab\t", "w" * 300]
    for it in range(40):
        ncols = rng.randint(1, 6)
        types = [rng.choice([T_I64, T_BOOL, T_STR, T_STR]) for _ in range(ncols)]
        n = rng.choice([0, 1, 33, 1000, 20000])
        vals = []
        for t in types:
            if t == T_I64:
                vals.append([rng.choice([0, -1, 9, 10, -10, 2**63 - 1, -2**63, rng.randint(-10**18, 10**18)]) for _ in range(n)])
            elif t == T_BOOL:
                vals.append([rng.random() < 0.5 for _ in range(n)])
            else:
                vals.append([rng.choice(words) for _ in range(n)])
        cols = [Column.from_values(v, t) for v, t in zip(vals, types)]
        sc = frontend.StageCompiler(types, [None] * ncols)
        sc.add_map(lambda x: x, 100001)
        st = backend.Stage(sc.finish_memory())
        res = st.run_host(0, cols, n)
        d = rng.choice([",", ";", "|"])
        assert res.csv_bytes(delimiter=d) == po.csv_write(cols, n, delimiter=d), (it, types, n)
        if ncols > 1:
            assert res.csv_bytes(delimiter=d, n_cols=ncols - 1) == po.csv_write(cols[:-1], n, delimiter=d)
        res.free()
        st.close()


def test_sink_f64_columns(gpu):
    """f64 cells: 8 fixed decimals, correctly rounded (ryu d2fixed(8) digits) — exact on the device below 2^63"""
    from tuplex_b200.backend import Column
    rng = random.Random(3)
    vals = [0.0, -0.0, 1.5, 2**-9, 0.999999995, 1 / 3, 1801.0, 123456789.123456789, -2.5e-9, float("inf"), float("-inf"), float("nan"), 2.0**62]
    vals += [rng.uniform(-1e7, 1e7) for _ in range(5000)] + [rng.randint(0, 2**20) * 2.0 ** -rng.randint(1, 40) for _ in range(5000)]
    cols = [Column.from_values(vals, T_F64), Column.from_values(list(range(len(vals))), T_I64)]
    sc = frontend.StageCompiler([T_F64, T_I64], [None, None])
    sc.add_map(lambda x: x, 100001)
    st = backend.Stage(sc.finish_memory())
    res = st.run_host(0, cols, len(vals))
    assert res.csv_bytes() == po.csv_write(cols, len(vals))
    res.free()
    res = st.run_host(0, [Column.from_values([1.0, 1e30], T_F64), Column.from_values([1, 2], T_I64)], 2)
    assert res.csv_bytes() is None  # magnitude >= 2^63: the call reports TPLX_E_UNSUPPORTED, tocsv formats on the host


# ---- lazy (late materialised) string columns -------------------------------------------------------------------------------
def test_lazy_columns_zillow_md5(gpu):
    """string columns the Z1 prefilter does not read stay as cell references in the CSV buffer and are gathered for the
    surviving rows only; the result must not change"""
    from tuplex_b200.dataset import csv_lazy_columns
    data = _zillow_csv()
    prog = workloads.zillow_program()
    lazy = csv_lazy_columns(prog, [c for c, t in enumerate(ZILLOW_FILE_TYPES) if t != X], ZILLOW_FILE_TYPES)
    assert lazy and 0 not in lazy and 6 not in lazy  # title and "facts and features" feed the prefilter
    buf = backend.CsvBuffer(0, data)
    p = buf.parse(ZILLOW_FILE_TYPES, header=True, lazy=lazy)
    st = backend.Stage(prog)
    res = st.run(p.block)
    txt = workloads.rows_to_csv([c.to_values() for c in res.columns()], workloads.ZILLOW_OUT)
    assert hashlib.md5(txt).hexdigest() == "4d5ca0263b1a5058341a369116dee83a"
    # a stage without a prefilter cannot read such a block: loud error, no silent garbage
    sc = frontend.StageCompiler([t for t in ZILLOW_FILE_TYPES if t != X], [None] * 8)
    sc.add_map(lambda x: x, 100001)
    with pytest.raises(backend.GpuBackendError):
        backend.Stage(sc.finish_memory()).run(p.block)


def test_lazy_columns_with_escaped_cells_equal_oracle(gpu):
    from tuplex_b200.backend import Column
    from tuplex_b200.dataset import csv_lazy_columns
    from helpers import assert_result_equals_oracle
    rng = random.Random(41)
    words = ["plain", "x,y", 'say "hi"', "", "line\nbreak", 'q""q', "tail\"", "é é"]
    lines = []
    for i in range(20000):
        s1 = rng.choice(["keep me", "drop", "kkk", "nothing"])
        s2 = rng.choice(words)
        s3 = rng.choice(words)
        q = lambda v: '"' + v.replace('"', '""') + '"' if any(ch in v for ch in ',"\n') or rng.random() < 0.2 else v
        lines.append(f"{i},{q(s1)},{q(s2)},{q(s3)}")
    data = ("\n".join(lines) + "\n").encode()
    types = [I, S, S, S]
    sc = frontend.StageCompiler(types, ["i", "s1", "s2", "s3"])
    sc.add_filter(lambda x: 'k' in x['s1'] and x['s1'].find('e') >= 0, 100001)
    sc.add_with_column("t", lambda x: x['s2'].replace('a', 'b') + '|' + x['s3'].upper() + '|' + x['s2'].lower(), 100002)
    sc.add_with_column("u", lambda x: x['s3'].replace('"', "'") + x['s2'][1:], 100003)
    prog = sc.finish_memory()
    assert prog.prefilter is not None
    lazy = csv_lazy_columns(prog, [0, 1, 2, 3], types)
    assert lazy == [2, 3]
    buf = backend.CsvBuffer(0, data)
    p = buf.parse(types, lazy=lazy)
    res = backend.Stage(prog).run(p.block)
    o = po.csv_parse(data, types)
    cols = [Column(T_I64, o.columns[0])] + [Column(T_STR, np.frombuffer(b, np.uint8), off) for b, off in o.columns[1:]]
    ora = po.run_program(prog, cols, o.n_normal)
    assert ora.n_out > 1000
    assert_result_equals_oracle(res, ora, "lazy csv columns")
