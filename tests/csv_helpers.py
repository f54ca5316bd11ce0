"""Shared helpers for the CSV source tests: seeded CSV generators and the host-shim loader."""
import ctypes as ct
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
T_I64, T_F64, T_BOOL, T_STR, T_SKIP = 0, 1, 2, 3, 0xFF


def gen_csv(rng: random.Random, n_rows: int, types, dirty: float = 0.05, weird_quotes: float = 0.0, delimiter=",") -> bytes:
    """Rows of the given column types; `dirty` = probability of a cell that does not fit its type / of a wrong cell
    count; `weird_quotes` = probability of stray quote characters (exercises the sequential repair path)."""
    words = ["", "a", "house", "New York", "x,y", 'say "hi"', "line\nbreak", " pad ", "NULL", "3 bds , 2 ba", "é", "tab\there"]
    out = []
    for _ in range(n_rows):
        cells = []
        for t in types:
            bad = rng.random() < dirty
            if t == T_I64 or (t == T_SKIP and rng.random() < 0.3):
                v = str(rng.randint(-10**rng.randint(0, 18), 10**rng.randint(0, 18)))
                if rng.random() < 0.1:
                    v = " " + v + "\t"
                if bad:
                    v = rng.choice(["", "12a", "1.5", "--3", "+7", "-", "9" * 25])
            elif t == T_F64:
                v = rng.choice([f"{rng.uniform(-1e6, 1e6):.{rng.randint(0, 9)}f}", f"{rng.randint(0, 99999)}.{rng.randint(0, 99):02d}",
                                f"{rng.uniform(0, 10):.4e}", str(rng.randint(-1000, 1000)), "nan", "inf", "Infinity", "1e400", ".5", "5.",
                                "+3.25", "-0.0", "1801.0"])
                if bad:
                    v = rng.choice(["", "1.2.3", "abc", "-inf", "n", "infi", "1e", "$5", "0x10"])
            elif t == T_BOOL:
                v = rng.choice(["true", "False", "T", "f", "YES", "no", "y", "N"])
                if bad:
                    v = rng.choice(["", "1", "0", "tru", "yess", " true"])
            else:
                v = rng.choice(words)
                if bad and t == T_STR:
                    v = rng.choice(["", "plain"])
            q = rng.random()
            needs = any(ch in v for ch in (delimiter, '"', "\n", "\r"))
            if needs or q < 0.15:
                v = '"' + v.replace('"', '""') + '"'
            if rng.random() < weird_quotes:
                v = rng.choice(['a"b', '"ab"c', '"', 'x""', '"a" '])
            cells.append(v)
        if rng.random() < dirty / 2:
            if rng.random() < 0.5 and len(cells) > 1:
                cells.pop()
            else:
                cells.append("extra")
        out.append(delimiter.join(cells))
    nl = rng.choice(["\n", "\r\n", "\n\n"])
    txt = nl.join(out)
    if rng.random() < 0.7:
        txt += nl
    return txt.encode("utf-8")


_SO = None


def host_shim(tmpdir=None):
    global _SO
    if _SO is None:
        so = os.path.join(tmpdir or "/tmp", f"csv_host_{os.getpid()}.so")
        subprocess.check_call(["g++", "-O1", "-x", "c++", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                               os.path.join(HERE, "csv_host.cpp")])
        L = ct.CDLL(so)
        L.hcsv_parse.restype = ct.c_void_p
        L.hcsv_parse.argtypes = [ct.c_char_p, ct.c_uint32, ct.c_uint8, ct.c_uint8, ct.c_int, ct.c_uint32, ct.c_char_p, ct.c_uint32,
                                 ct.POINTER(ct.c_char_p)]
        for f, rt in (("hcsv_fixed", ct.POINTER(ct.c_uint64)), ("hcsv_offsets", ct.POINTER(ct.c_uint32)),
                      ("hcsv_bytes", ct.POINTER(ct.c_uint8)), ("hcsv_rowmap", ct.POINTER(ct.c_uint32)), ("hcsv_bad", ct.POINTER(ct.c_uint32))):
            getattr(L, f).restype = rt
        for f in ("hcsv_counts", "hcsv_type", "hcsv_fixed", "hcsv_offsets", "hcsv_rowmap", "hcsv_bad", "hcsv_free"):
            getattr(L, f).argtypes = [ct.c_void_p] + ([ct.c_uint32] if f in ("hcsv_type", "hcsv_fixed", "hcsv_offsets") else []) + \
                                     ([ct.POINTER(ct.c_uint32)] if f == "hcsv_counts" else [])
        L.hcsv_bytes.argtypes = [ct.c_void_p, ct.c_uint32, ct.POINTER(ct.c_uint64)]
        L.hcsv_write.restype = ct.c_ulonglong
        L.hcsv_write.argtypes = [ct.c_uint, ct.c_char_p, ct.POINTER(ct.c_void_p), ct.POINTER(ct.c_void_p), ct.POINTER(ct.c_void_p),
                                 ct.c_ulonglong, ct.c_uint8, ct.c_uint8, ct.c_void_p]
        L.hcsv_atod.argtypes = [ct.c_char_p, ct.c_uint32, ct.POINTER(ct.c_double)]
        L.hcsv_atob.argtypes = [ct.c_char_p, ct.c_uint32, ct.POINTER(ct.c_longlong)]
        _SO = L
    return _SO


class Parsed:
    """n_rows, columns (np arrays or (bytes, offsets)), rowmap, bad [(row, code, start, end)], sequential flag"""

    def __init__(self, n_rows, columns, types, rowmap, bad, sequential=0):
        self.n_rows, self.columns, self.types, self.rowmap, self.bad, self.sequential = n_rows, columns, types, rowmap, bad, sequential


def host_parse(data: bytes, col_types, delimiter=",", quotechar='"', header=False, null_values=("",)) -> Parsed:
    L = host_shim()
    nulls = (ct.c_char_p * max(1, len(null_values)))(*[s.encode() for s in null_values])
    h = L.hcsv_parse(data, len(data), ord(delimiter), ord(quotechar), int(header), len(col_types), bytes(col_types), len(null_values), nulls)
    cnt = (ct.c_uint32 * 5)()
    L.hcsv_counts(h, cnt)
    n_out, n_rows, n_good, n_bad, seq = list(cnt)
    cols, types = [], []
    for c in range(n_out):
        t = L.hcsv_type(h, c)
        types.append(t)
        if t == T_STR:
            offs = np.ctypeslib.as_array(L.hcsv_offsets(h, c), (n_good + 1,)).copy()
            nb = ct.c_uint64()
            p = L.hcsv_bytes(h, c, ct.byref(nb))
            cols.append((ct.string_at(p, nb.value) if nb.value else b"", offs))
        else:
            a = np.ctypeslib.as_array(L.hcsv_fixed(h, c), (n_good,)).copy().view(np.int64) if n_good else np.zeros(0, np.int64)
            cols.append(a.view(np.float64) if t == T_F64 else a)
    rowmap = np.ctypeslib.as_array(L.hcsv_rowmap(h), (n_good,)).copy() if n_good else np.zeros(0, np.uint32)
    badarr = np.ctypeslib.as_array(L.hcsv_bad(h), (n_bad * 4,)).copy().reshape(-1, 4) if n_bad else np.zeros((0, 4), np.uint32)
    L.hcsv_free(h)
    return Parsed(n_rows, cols, types, rowmap, [tuple(int(x) for x in r) for r in badarr], seq)


def assert_same_parse(a, b, what=""):
    """a, b: Parsed / CsvOracleResult. Floats are compared by bit pattern (NaN == NaN)."""
    assert a.n_rows == b.n_rows, (what, a.n_rows, b.n_rows)
    assert list(a.types) == list(b.types), what
    assert np.array_equal(np.asarray(a.rowmap, np.uint32), np.asarray(b.rowmap, np.uint32)), what
    assert [tuple(x) for x in a.bad] == [tuple(x) for x in b.bad], (what, a.bad[:5], b.bad[:5])
    for c, (x, y) in enumerate(zip(a.columns, b.columns)):
        if isinstance(x, tuple):
            assert x[0] == y[0], (what, "string bytes of column", c)
            assert np.array_equal(x[1], y[1]), (what, "offsets of column", c)
        else:
            assert np.array_equal(np.asarray(x).view(np.int64), np.asarray(y).view(np.int64)), (what, "values of column", c)


def host_csv_write(cols, n_rows: int, delimiter=",", quotechar='"') -> bytes:
    """The device row writer (csvops.cuh csv_sink_*) run on the host over Column-like objects."""
    L = host_shim()
    n = len(cols)
    keep = []
    data = (ct.c_void_p * n)()
    offs = (ct.c_void_p * n)()
    byts = (ct.c_void_p * n)()
    for i, c in enumerate(cols):
        d = np.ascontiguousarray(c.data)
        keep.append(d)
        if c.type == T_STR:
            o = np.ascontiguousarray(c.offsets, dtype=np.uint32)
            keep.append(o)
            offs[i] = o.ctypes.data
            byts[i] = d.ctypes.data
        else:
            data[i] = d.ctypes.data
    types = bytes(c.type for c in cols)
    need = L.hcsv_write(n, types, data, offs, byts, n_rows, ord(delimiter), ord(quotechar), None)
    buf = np.empty(need + 1, dtype=np.uint8)
    L.hcsv_write(n, types, data, offs, byts, n_rows, ord(delimiter), ord(quotechar), buf.ctypes.data)
    return buf[:need].tobytes()
