"""CPU ORACLE — Python side. TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/liboracle.so (tplx_oracle.c / workloads.c) plus a tiny pure-CPython
row evaluator used to cross-check the UDF front end on inputs where Tuplex and CPython agree.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under tuplex_b200/ does.

Reference citations live in tplx_oracle.c. Parity pinning: oracle/README.md.
"""
from __future__ import annotations

import ctypes as ct
import os
import struct
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
MAX_COLS, MAX_ACCS = 64, 16
T_I64, T_F64, T_BOOL, T_STR = 0, 1, 2, 3


class OCol(ct.Structure):
    _fields_ = [("type", ct.c_uint8), ("pad", ct.c_uint8 * 7), ("data", ct.c_void_p), ("offsets", ct.c_void_p),
                ("data_bytes", ct.c_uint64), ("valid", ct.c_void_p)]


class OExc(ct.Structure):
    _fields_ = [("row", ct.c_int64), ("row_no", ct.c_int64), ("code", ct.c_int64), ("op_id", ct.c_int64)]


class OResult(ct.Structure):
    _fields_ = [("n_out", ct.c_uint64), ("n_exc", ct.c_uint64), ("n_accs", ct.c_uint64), ("n_cols", ct.c_uint64),
                ("col_types", ct.c_uint8 * MAX_COLS),
                ("fixed", ct.c_void_p * MAX_COLS), ("offsets", ct.c_void_p * MAX_COLS), ("bytes", ct.c_void_p * MAX_COLS),
                ("bytes_len", ct.c_uint64 * MAX_COLS), ("bytes_cap", ct.c_uint64 * MAX_COLS),
                ("exc", ct.c_void_p), ("exc_cap", ct.c_uint64), ("out_cap", ct.c_uint64),
                ("acc_seq", ct.c_int64 * MAX_ACCS), ("acc_tree", ct.c_int64 * MAX_ACCS)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ct.CDLL(_LIB)
        L.tplx_oracle_run.argtypes = [ct.c_void_p, ct.c_uint64, ct.POINTER(OCol), ct.c_uint64, ct.c_int64, ct.c_uint32,
                                      ct.c_uint32, ct.c_uint32, ct.POINTER(OResult)]
        L.tplx_oracle_run.restype = ct.c_int
        L.tplx_oracle_free.argtypes = [ct.POINTER(OResult)]
        L.tplx_oracle_to_partitions.argtypes = [ct.POINTER(OCol), ct.c_uint32, ct.c_uint64, ct.c_uint64, ct.c_void_p,
                                                ct.POINTER(ct.c_uint64), ct.c_uint32, ct.POINTER(ct.c_uint32)]
        L.tplx_oracle_to_partitions.restype = ct.c_uint64
        L.tplx_oracle_exception_partition.argtypes = [ct.POINTER(OCol), ct.c_uint32, ct.c_void_p, ct.c_uint64, ct.c_void_p]
        L.tplx_oracle_exception_partition.restype = ct.c_uint64
        L.tplx_oracle_q6.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_uint64, ct.c_uint64, ct.c_int]
        L.tplx_oracle_q6.restype = ct.c_double
        L.tplx_oracle_c1.argtypes = [ct.c_void_p, ct.c_uint64, ct.c_void_p, ct.c_uint64, ct.c_int]
        L.tplx_oracle_c1.restype = ct.c_uint64
        L.tplx_o_atoi64.argtypes = [ct.c_char_p, ct.c_int64, ct.POINTER(ct.c_int64)]
        L.tplx_o_atoi64.restype = ct.c_int32
        L.tplx_o_floordiv.argtypes = [ct.c_int64, ct.c_int64]
        L.tplx_o_floordiv.restype = ct.c_int64
        L.tplx_o_floormod.argtypes = [ct.c_int64, ct.c_int64]
        L.tplx_o_floormod.restype = ct.c_int64
        _lib = L
    return _lib


def _ocols(cols):
    """cols: sequence of objects with .type, .data (np array), .offsets (np array or None)."""
    arr = (OCol * max(len(cols), 1))()
    keep = []
    for i, c in enumerate(cols):
        d = np.ascontiguousarray(c.data)
        keep.append(d)
        arr[i].type = c.type
        arr[i].data = d.ctypes.data
        if c.type == T_STR:
            o = np.ascontiguousarray(c.offsets, dtype=np.uint32)
            keep.append(o)
            arr[i].offsets = o.ctypes.data
            arr[i].data_bytes = int(o[-1]) if len(o) else 0
        else:
            arr[i].data_bytes = d.nbytes
        v = getattr(c, "valid", None)
        if v is not None:  # Option[T] column (row formats: bitmap; the op program reads companions instead)
            v = np.ascontiguousarray(v, dtype=np.uint32)
            keep.append(v)
            arr[i].valid = v.ctypes.data
    return arr, keep


class OracleResult:
    def __init__(self):
        self.n_out = 0
        self.columns: List[Tuple[int, np.ndarray, Optional[np.ndarray]]] = []  # (type, data, offsets)
        self.exceptions = np.zeros(0, dtype=[("row", "<i8"), ("row_no", "<i8"), ("code", "<i8"), ("op_id", "<i8")])
        self.acc_seq: List[int] = []
        self.acc_tree: List[int] = []

    def values(self, c: int) -> list:
        t, data, offs = self.columns[c]
        if t == T_STR:
            raw = data.tobytes()
            return [raw[offs[i]:offs[i + 1]].decode("utf-8") for i in range(self.n_out)]
        if t == T_F64:
            return data.view(np.float64).tolist()
        if t == T_BOOL:
            return [bool(v) for v in data.tolist()]
        return data.tolist()


def run_program(program, cols, n_rows: int, first_row_no: int = 0, tile_R: int = 16, tile_NT: int = 256,
                fin_NT: int = 1024) -> OracleResult:
    """Run a stage descriptor (tuplex_b200.ir.Program or its serialized bytes) over host columns."""
    blob = program if isinstance(program, (bytes, bytearray)) else program.serialize()
    buf = ct.create_string_buffer(bytes(blob), len(blob))
    arr, keep = _ocols(cols)
    res = OResult()
    rc = lib().tplx_oracle_run(buf, len(blob), arr, n_rows, first_row_no, tile_R, tile_NT, fin_NT, ct.byref(res))
    if rc != 0:
        raise RuntimeError("oracle rejected the stage descriptor")
    out = OracleResult()
    out.n_out = int(res.n_out)
    n = out.n_out
    is_agg = res.n_accs > 0 and res.n_cols == 0
    if not is_agg:
        for c in range(res.n_cols):
            t = res.col_types[c]
            if t == T_STR:
                offs = np.ctypeslib.as_array(ct.cast(res.offsets[c], ct.POINTER(ct.c_uint32)), shape=(n + 1,)).copy()
                nb = int(res.bytes_len[c])
                data = np.ctypeslib.as_array(ct.cast(res.bytes[c], ct.POINTER(ct.c_uint8)), shape=(nb,)).copy() if nb else np.zeros(0, np.uint8)
                out.columns.append((t, data, offs))
            else:
                data = np.ctypeslib.as_array(ct.cast(res.fixed[c], ct.POINTER(ct.c_int64)), shape=(n,)).copy() if n else np.zeros(0, np.int64)
                out.columns.append((t, data, None))
    ne = int(res.n_exc)
    if ne:
        raw = ct.string_at(res.exc, ne * 32)
        out.exceptions = np.frombuffer(raw, dtype=out.exceptions.dtype).copy()
    out.acc_seq = [res.acc_seq[k] & ((1 << 64) - 1) for k in range(res.n_accs)]
    out.acc_tree = [res.acc_tree[k] & ((1 << 64) - 1) for k in range(res.n_accs)]
    lib().tplx_oracle_free(ct.byref(res))
    return out


def to_partitions(cols, n_rows: int, partition_bytes: int = 32 << 20) -> List[bytes]:
    arr, keep = _ocols(cols)
    np_ = ct.c_uint32()
    total = lib().tplx_oracle_to_partitions(arr, len(cols), n_rows, partition_bytes, None, None, 0, ct.byref(np_))
    buf = np.zeros(max(total, 8), dtype=np.uint8)
    offs = (ct.c_uint64 * (np_.value + 1))()
    lib().tplx_oracle_to_partitions(arr, len(cols), n_rows, partition_bytes, buf.ctypes.data, offs, np_.value, ct.byref(np_))
    raw = buf.tobytes()
    return [raw[offs[p]:offs[p + 1]] for p in range(np_.value)]


def exception_partition(in_cols, exc: np.ndarray) -> bytes:
    arr, keep = _ocols(in_cols)
    e = np.ascontiguousarray(exc)
    total = lib().tplx_oracle_exception_partition(arr, len(in_cols), e.ctypes.data, len(e), None)
    buf = np.zeros(total, dtype=np.uint8)
    lib().tplx_oracle_exception_partition(arr, len(in_cols), e.ctypes.data, len(e), buf.ctypes.data)
    return buf.tobytes()


def q6(qty: np.ndarray, price: np.ndarray, disc: np.ndarray, ship: np.ndarray, part_rows: int = 0, threads: int = 1) -> float:
    qty, price, disc, ship = (np.ascontiguousarray(a) for a in (qty, price, disc, ship))
    return lib().tplx_oracle_q6(qty.ctypes.data, price.ctypes.data, disc.ctypes.data, ship.ctypes.data, len(qty), part_rows, threads)


def c1(x: np.ndarray, part_rows: int = 0, threads: int = 1) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.int64)
    out = np.empty(len(x), dtype=np.int64)
    n = lib().tplx_oracle_c1(x.ctypes.data, len(x), out.ctypes.data, part_rows, threads)
    return out[:n]


def atoi64(s: str):
    """int(str) exactly as the reference parses it; returns (ok, value)."""
    b = s.encode()
    v = ct.c_int64()
    rc = lib().tplx_o_atoi64(b, len(b), ct.byref(v))
    return rc == 0, v.value


# ---- pure CPython evaluation of a pipeline (small cases only) -----------------------------------------
def cpython_pipeline(rows: Sequence, ops: Sequence[Tuple[str, object]]):
    """Evaluate [('map', f), ('filter', f), ...] row by row in CPython. Rows whose UDF raises are
    returned separately as (index, exception type name). Used to cross-check the front end on inputs
    where the reference's semantics and CPython's coincide."""
    out, bad = [], []
    for i, r in enumerate(rows):
        try:
            keep = True
            for kind, f in ops:
                if kind == "map":
                    r = f(r)
                elif kind == "filter":
                    if not f(r):
                        keep = False
                        break
            if keep:
                out.append(r)
        except Exception as e:  # noqa: BLE001
            bad.append((i, type(e).__name__))
    return out, bad


# ---------------------------------------------------------------------------------------------
# CSV source oracle (csv_oracle.c)
# ---------------------------------------------------------------------------------------------
T_SKIP = 0xFF


class _CsvBad(ct.Structure):
    _fields_ = [("row", ct.c_uint32), ("code", ct.c_uint32), ("line_start", ct.c_uint32), ("line_end", ct.c_uint32)]


class _CsvResult(ct.Structure):
    _fields_ = [("n_out_cols", ct.c_uint32), ("out_types", ct.c_uint8 * 256),
                ("n_rows", ct.c_uint64), ("n_normal", ct.c_uint64), ("n_bad", ct.c_uint64),
                ("fixed", ct.c_void_p * 256), ("offsets", ct.c_void_p * 256), ("bytes", ct.c_void_p * 256),
                ("bytes_len", ct.c_uint64 * 256), ("bytes_cap", ct.c_uint64 * 256),
                ("rowmap", ct.c_void_p), ("bad", ct.c_void_p), ("cap_rows", ct.c_uint64), ("cap_bad", ct.c_uint64),
                ("dump", ct.c_void_p), ("dump_len", ct.c_uint64), ("dump_cap", ct.c_uint64)]


class CsvOracleResult:
    """columns: list of np.int64/float64/bool arrays or (bytes, offsets) pairs for the parsed (non-skipped) columns."""

    def __init__(self, n_rows, columns, types, rowmap, bad, dump):
        self.n_rows, self.columns, self.types, self.rowmap, self.bad, self.dump = n_rows, columns, types, rowmap, bad, dump
        self.n_normal = len(rowmap)


def csv_parse(data: bytes, col_types: Sequence[int], delimiter=",", quotechar='"', header=False,
              null_values: Sequence[str] = ("",), dump_cells=False) -> CsvOracleResult:
    L = lib()
    L.csv_oracle_parse.restype = ct.POINTER(_CsvResult)
    L.csv_oracle_parse.argtypes = [ct.c_char_p, ct.c_uint64, ct.c_char, ct.c_char, ct.c_int, ct.c_uint32, ct.c_char_p,
                                   ct.c_uint32, ct.POINTER(ct.c_char_p), ct.c_int]
    L.csv_oracle_free.argtypes = [ct.POINTER(_CsvResult)]
    nulls = (ct.c_char_p * max(1, len(null_values)))(*[s.encode() for s in null_values])
    rp = L.csv_oracle_parse(data, len(data), delimiter.encode(), quotechar.encode(), int(header), len(col_types),
                            bytes(col_types), len(null_values), nulls, int(dump_cells))
    r = rp.contents
    n = int(r.n_normal)
    cols, types = [], []
    for c in range(r.n_out_cols):
        t = r.out_types[c]
        types.append(t)
        if t == T_STR:
            offs = np.ctypeslib.as_array(ct.cast(r.offsets[c], ct.POINTER(ct.c_uint32)), (n + 1,)).copy()
            nb = int(r.bytes_len[c])
            by = ct.string_at(r.bytes[c], nb) if nb else b""
            cols.append((by, offs))
        else:
            a = np.ctypeslib.as_array(ct.cast(r.fixed[c], ct.POINTER(ct.c_int64)), (n,)).copy() if n else np.zeros(0, np.int64)
            cols.append(a.view(np.float64) if t == T_F64 else a)
    rowmap = np.ctypeslib.as_array(ct.cast(r.rowmap, ct.POINTER(ct.c_uint32)), (n,)).copy() if n else np.zeros(0, np.uint32)
    bad = [(b.row, b.code, b.line_start, b.line_end) for b in
           (ct.cast(r.bad, ct.POINTER(_CsvBad))[i] for i in range(int(r.n_bad)))]
    dump = ct.string_at(r.dump, int(r.dump_len)) if r.dump_len else b""
    out = CsvOracleResult(int(r.n_rows), cols, types, rowmap, bad, dump)
    L.csv_oracle_free(rp)
    return out


def csv_ref_cells(data: bytes, delimiter=",", quotechar='"') -> Optional[bytes]:
    """Cell dump of the reference's own csvmonkey reader (oracle/_ref/csv_ref); None when the binary is absent."""
    exe = os.path.join(_HERE, "_ref", "csv_ref")
    if not os.path.exists(exe):
        return None
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".csv") as f:
        f.write(data)
        f.flush()
        return subprocess.run([exe, f.name, delimiter, quotechar], check=True, capture_output=True).stdout


def csv_scalar(kind: str, s: str):
    """fast_atoi64 / fast_atod / fast_atob behind the runtime's trimming wrapper; None = ValueError."""
    L = lib()
    if kind == "i64":
        v = ct.c_int64()
        return None if L.csv_oracle_atoi64(s.encode("latin1"), ct.byref(v)) else v.value
    if kind == "f64":
        d = ct.c_double()
        return None if L.csv_oracle_atod(s.encode("latin1"), ct.byref(d)) else d.value
    b = ct.c_int()
    return None if L.csv_oracle_atob(s.encode("latin1"), ct.byref(b)) else bool(b.value)


def csv_write(cols, n_rows: int, delimiter=",", quotechar='"') -> bytes:
    """CSV sink oracle (fast_csvwriter + quoteForCSV): cols = backend.Column-like objects of type i64 / bool / str."""
    L = lib()
    arr, keep = _ocols(cols)
    L.csv_oracle_write.restype = ct.c_uint64
    L.csv_oracle_write.argtypes = [ct.c_void_p, ct.c_uint32, ct.c_uint64, ct.c_char, ct.c_char, ct.c_void_p]
    need = L.csv_oracle_write(arr, len(cols), n_rows, delimiter.encode(), quotechar.encode(), None)
    buf = np.empty(need, dtype=np.uint8)
    L.csv_oracle_write(arr, len(cols), n_rows, delimiter.encode(), quotechar.encode(), buf.ctypes.data if need else None)
    return buf.tobytes()


# ---- hash join (join_oracle.c) ----------------------------------------------------------------------------------------------
class _JCol(ct.Structure):
    _fields_ = [("type", ct.c_uint8), ("pad", ct.c_uint8 * 7), ("data", ct.c_void_p), ("offsets", ct.c_void_p), ("valid", ct.c_void_p)]


def _jcol(col):
    """col: object with .type, .data, .offsets and optionally .valid (uint32 words, bit set = present)."""
    jc = _JCol()
    keep = []
    d = np.ascontiguousarray(col.data)
    keep.append(d)
    jc.type = col.type
    jc.data = d.ctypes.data
    if col.type == T_STR:
        o = np.ascontiguousarray(col.offsets, dtype=np.uint32)
        keep.append(o)
        jc.offsets = o.ctypes.data
    v = getattr(col, "valid", None)
    if v is not None:
        v = np.ascontiguousarray(v, dtype=np.uint32)
        keep.append(v)
        jc.valid = v.ctypes.data
    return jc, keep


def join_pairs(build_key, n_build: int, probe_key, n_probe: int, left_outer: bool = False):
    """(probe rows, build rows) of the join's output in the reference's order; build row -1 = left join without a match."""
    L = lib()
    L.tplx_oracle_join.restype = ct.c_uint64
    L.tplx_oracle_join.argtypes = [ct.POINTER(_JCol), ct.c_uint64, ct.POINTER(_JCol), ct.c_uint64, ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_uint64]
    b, kb = _jcol(build_key)
    p, kp = _jcol(probe_key)
    n = L.tplx_oracle_join(ct.byref(b), n_build, ct.byref(p), n_probe, int(left_outer), None, None, 0)
    op = np.empty(n, np.int64)
    ob = np.empty(n, np.int64)
    L.tplx_oracle_join(ct.byref(b), n_build, ct.byref(p), n_probe, int(left_outer), op.ctypes.data if n else None, ob.ctypes.data if n else None, n)
    return op, ob


def join_rows(left_rows: Sequence[tuple], left_key: int, right_rows: Sequence[tuple], right_key: int, key_type: int, left_outer: bool = False,
              build_right: bool = True) -> List[tuple]:
    """The join of two lists of python tuples through the C oracle (None keys allowed): the rows the reference's collect() returns,
    | left non-key | key | right non-key | (JoinOperator.cc:163-184), ordered by the probe side."""
    class _K:
        pass

    def keycol(rows, k):
        c = _K()
        c.type = key_type
        vals = [r[k] for r in rows]
        valid = np.zeros((len(vals) + 31) // 32, np.uint32)
        for i, v in enumerate(vals):
            if v is not None:
                valid[i >> 5] |= np.uint32(1 << (i & 31))
        c.valid = valid if any(v is None for v in vals) else None
        if key_type == T_STR:
            enc = [(v or "").encode() for v in vals]
            c.offsets = np.zeros(len(enc) + 1, np.uint32)
            np.cumsum([len(e) for e in enc], out=c.offsets[1:])
            c.data = np.frombuffer(b"".join(enc) + b"\0", np.uint8).copy()
        else:
            c.offsets = None
            c.data = np.array([int(v or 0) for v in vals], np.int64)
        return c

    if build_right:
        op, ob = join_pairs(keycol(right_rows, right_key), len(right_rows), keycol(left_rows, left_key), len(left_rows), left_outer)
        li, ri = op, ob
    else:
        assert not left_outer
        op, ob = join_pairs(keycol(left_rows, left_key), len(left_rows), keycol(right_rows, right_key), len(right_rows), False)
        li, ri = ob, op
    n_right = len(right_rows[0]) if right_rows else 0
    out = []
    for l, r in zip(li.tolist(), ri.tolist()):
        lrow = left_rows[l]
        rrow = right_rows[r] if r >= 0 else (None,) * n_right
        key = lrow[left_key] if build_right else right_rows[r][right_key]
        out.append(tuple(v for i, v in enumerate(lrow) if i != left_key) + (key,) + tuple(v for i, v in enumerate(rrow) if i != right_key))
    return out


# ---- in-order merge of resolved rows (merge_oracle.c) ------------------------------------------------------------------------
def merge_sources(n_norm: int, exc_row_nos: Sequence[int], resolved: Sequence[bool]) -> np.ndarray:
    """Source of every row of the merged output stream: j >= 0 = normal row j, ~m = the m-th resolved exception (ResolveTask::executeInOrder)."""
    L = lib()
    L.tplx_oracle_merge.restype = ct.c_uint64
    L.tplx_oracle_merge.argtypes = [ct.c_uint64, ct.c_void_p, ct.c_void_p, ct.c_uint64, ct.c_void_p]
    nos = np.ascontiguousarray(exc_row_nos, dtype=np.int64)
    res = np.ascontiguousarray(resolved, dtype=np.uint8)
    out = np.empty(n_norm + len(nos) + 1, np.int64)
    n = L.tplx_oracle_merge(n_norm, nos.ctypes.data, res.ctypes.data, len(nos), out.ctypes.data)
    return out[:n].copy()
