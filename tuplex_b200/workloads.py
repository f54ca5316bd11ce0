"""The benchmark pipelines of BASELINE.json, expressed against this package's API, and their
synthetic-input generators.

  C1  parallelize([1..1e6]).map(x*x).filter(x%2==0)
  C2  Zillow Z1 cleaning pipeline          workload spec: benchmarks/zillow/Z1/runtuplex.py:12-205
  C3  TPC-H Q6 (pre-processed columns)      workload spec: benchmarks/tpch/Q06/runtuplex.py:96-99
  C5  aggregateByKey, string keys           tuplex/test/core/AggregateTest.cc:249-364 shape

The UDFs below are the user code of those workloads (same operations on the same columns, so that the
output is byte-identical to the reference baselines'), not library code.
"""
from __future__ import annotations

import csv
import gzip
import io
import os
from typing import List, Optional, Tuple

import numpy as np

from .backend import Column
from .frontend import StageCompiler
from .ir import Program, T_F64, T_I64, T_STR

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
ZILLOW_COLS = ["title", "address", "city", "state", "postal_code", "price", "facts and features", "url"]
ZILLOW_TYPES = [T_STR, T_STR, T_STR, T_STR, T_F64, T_STR, T_STR, T_STR]
ZILLOW_OUT = ["url", "zipcode", "address", "city", "state", "bedrooms", "bathrooms", "sqft", "offer", "type", "price"]


# ---- C2: Zillow ----------------------------------------------------------------------------------------
def _number_before(x, marker):
    # "<...>, 3 bds , 2 ba , 1,560 sqft": the number in front of `marker`, after the previous comma
    ff = x['facts and features']
    stop = ff.find(marker)
    if stop < 0:
        stop = len(ff)
    head = ff[:stop]
    start = head.rfind(',')
    if start < 0:
        start = 0
    else:
        start += 2
    return int(head[start:])


def bedrooms_of(x):
    ff = x['facts and features']
    stop = ff.find(' bd')
    if stop < 0:
        stop = len(ff)
    head = ff[:stop]
    start = head.rfind(',')
    if start < 0:
        start = 0
    else:
        start += 2
    return int(head[start:])


def bathrooms_of(x):
    ff = x['facts and features']
    stop = ff.find(' ba')
    if stop < 0:
        stop = len(ff)
    head = ff[:stop]
    start = head.rfind(',')
    if start < 0:
        start = 0
    else:
        start += 2
    return int(head[start:])


def sqft_of(x):
    ff = x['facts and features']
    stop = ff.find(' sqft')
    if stop < 0:
        stop = len(ff)
    head = ff[:stop]
    start = head.rfind('ba ,')
    if start < 0:
        start = 0
    else:
        start += 5
    digits = head[start:].replace(',', '')
    return int(digits)


def offer_of(x):
    t = x['title'].lower()
    if 'sale' in t:
        return 'sale'
    if 'rent' in t:
        return 'rent'
    if 'sold' in t:
        return 'sold'
    if 'foreclose' in t.lower():
        return 'foreclosed'
    return t


def type_of(x):
    t = x['title'].lower()
    kind = 'unknown'
    if 'condo' in t or 'apartment' in t:
        kind = 'condo'
    if 'house' in t:
        kind = 'house'
    return kind


def price_of(x):
    raw = x['price']
    value = 0
    if x['offer'] == 'sold':
        # sold listings carry price/sqft in the facts column: price = price/sqft * sqft
        ff = x['facts and features']
        tail = ff[ff.find('Price/sqft:') + len('Price/sqft:') + 1:]
        per_sqft = int(tail[tail.find('$') + 1:tail.find(', ') - 1])
        value = per_sqft * x['sqft']
    elif x['offer'] == 'rent':
        slash = raw.rfind('/')
        value = int(raw[1:slash].replace(',', ''))
    else:
        value = int(raw[1:].replace(',', ''))
    return value


def zillow_pipeline(ds):
    """Z1 operator chain (benchmarks/zillow/Z1/runtuplex.py:192-205) on a DataSet with ZILLOW_COLS."""
    return (ds.withColumn("bedrooms", bedrooms_of)
            .filter(lambda x: x['bedrooms'] < 10)
            .withColumn("type", type_of)
            .filter(lambda x: x['type'] == 'house')
            .withColumn("zipcode", lambda x: '%05d' % int(x['postal_code']))
            .mapColumn("city", lambda x: x[0].upper() + x[1:].lower())
            .withColumn("bathrooms", bathrooms_of)
            .withColumn("sqft", sqft_of)
            .withColumn("offer", offer_of)
            .withColumn("price", price_of)
            .filter(lambda x: 100000 < x['price'] < 2e7)
            .selectColumns(ZILLOW_OUT))


def zillow_program(first_op_id: int = 100001) -> Program:
    """The same chain lowered straight to a stage descriptor (what bench.py hands to the C ABI)."""
    sc = StageCompiler(ZILLOW_TYPES, ZILLOW_COLS)
    k = first_op_id
    sc.add_with_column("bedrooms", bedrooms_of, k)
    sc.add_filter(lambda x: x['bedrooms'] < 10, k + 1)
    sc.add_with_column("type", type_of, k + 2)
    sc.add_filter(lambda x: x['type'] == 'house', k + 3)
    sc.add_with_column("zipcode", lambda x: '%05d' % int(x['postal_code']), k + 4)
    sc.add_map_column("city", lambda x: x[0].upper() + x[1:].lower(), k + 5)
    sc.add_with_column("bathrooms", bathrooms_of, k + 6)
    sc.add_with_column("sqft", sqft_of, k + 7)
    sc.add_with_column("offer", offer_of, k + 8)
    sc.add_with_column("price", price_of, k + 9)
    sc.add_filter(lambda x: 100000 < x['price'] < 2e7, k + 10)
    sc.add_select(ZILLOW_OUT, k + 11)
    return sc.finish_memory()


def load_zillow_fixture() -> Tuple[List[Column], int]:
    """The 32,661 cleaned rows of the reference's zillow_noexc.csv (8 referenced columns)."""
    with gzip.open(os.path.join(GOLDEN, "zillow_noexc_cols.csv.gz"), "rt", newline="") as fp:
        rows = list(csv.reader(fp))[1:]
    cols = []
    for c, t in enumerate(ZILLOW_TYPES):
        if t == T_F64:
            cols.append(Column.from_values([float(r[c]) for r in rows], T_F64))
        else:
            cols.append(Column.from_values([r[c] for r in rows], T_STR))
    return cols, len(rows)


def zillow_golden_csv() -> bytes:
    with gzip.open(os.path.join(GOLDEN, "zillow_noexc_out.csv.gz"), "rb") as fp:
        return fp.read()


def replicate(cols: List[Column], n_src: int, n: int) -> List[Column]:
    """Cyclic replication (itertools.cycle, no shuffling) of the source rows up to n rows —
    the reference's own data generator, benchmarks/zillow/Z1/sample_zillow.py:20-45."""
    reps, rem = divmod(n, n_src)
    out = []
    for c in cols:
        if c.type == T_STR:
            o = c.offsets.astype(np.int64)
            total = int(o[-1])
            lens = np.diff(o)
            data = np.concatenate([np.tile(c.data[:total], reps), c.data[: int(o[rem])]]) if n else np.zeros(0, np.uint8)
            all_lens = np.concatenate([np.tile(lens, reps), lens[:rem]])
            offs = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(all_lens, out=offs[1:])
            if offs[-1] > 0xFFFFFFFF:
                raise ValueError("replicated string column exceeds 4 GiB: replicate per block instead")
            out.append(Column(T_STR, data, offs.astype(np.uint32)))
        else:
            out.append(Column(c.type, np.concatenate([np.tile(c.data, reps), c.data[:rem]])))
    return out


def rows_to_csv(columns_values: List[list], header: Optional[List[str]]) -> bytes:
    """Unquoted CSV, exactly what zillow.cpp's snprintf("%s,%s,...") writes."""
    buf = io.StringIO()
    if header:
        buf.write(",".join(header) + "\n")
    n = len(columns_values[0]) if columns_values else 0
    for i in range(n):
        buf.write(",".join(str(col[i]) for col in columns_values) + "\n")
    return buf.getvalue().encode()


# ---- C3: TPC-H Q6 ----------------------------------------------------------------------------------------
Q6_COLS = ["l_quantity", "l_extended_price", "l_discount", "l_shipdate"]
Q6_TYPES = [T_I64, T_F64, T_F64, T_I64]


def q6_pipeline(ds):
    return (ds.filter(lambda x: 19940101 <= x['l_shipdate'] < 19950101)
            .filter(lambda x: 0.05 <= x['l_discount'] <= 0.07)
            .filter(lambda x: x['l_quantity'] < 24)
            .aggregate(lambda a, b: a + b, lambda a, x: a + x[1] * x[2], 0.0))


def q6_program(first_op_id: int = 100001) -> Program:
    sc = StageCompiler(Q6_TYPES, Q6_COLS)
    sc.add_filter(lambda x: 19940101 <= x['l_shipdate'] < 19950101, first_op_id)
    sc.add_filter(lambda x: 0.05 <= x['l_discount'] <= 0.07, first_op_id + 1)
    sc.add_filter(lambda x: x['l_quantity'] < 24, first_op_id + 2)
    return sc.finish_aggregate(lambda a, x: a + x[1] * x[2], lambda a, b: a + b, 0.0, first_op_id + 3)


def load_lineitem_fixture() -> List[Column]:
    z = np.load(os.path.join(GOLDEN, "lineitem_q6.npz"))
    return [Column(T_I64, z["l_quantity"]), Column(T_F64, z["l_extendedprice"]), Column(T_F64, z["l_discount"]),
            Column(T_I64, z["l_shipdate"])]


_DAYS = None


def _ship_dates() -> np.ndarray:
    """All yyyymmdd between 1992-01-02 and 1998-12-01 (dbgen's l_shipdate range)."""
    global _DAYS
    if _DAYS is None:
        d = np.arange(np.datetime64("1992-01-02"), np.datetime64("1998-12-02"))
        y = d.astype("datetime64[Y]").astype(int) + 1970
        m = d.astype("datetime64[M]").astype(int) % 12 + 1
        dd = (d - d.astype("datetime64[M]")).astype(int) + 1
        _DAYS = (y * 10000 + m * 100 + dd).astype(np.int64)
    return _DAYS


def gen_lineitem(n: int, seed: int = 42) -> List[Column]:
    """Synthetic lineitem columns with dbgen's value ranges: quantity 1..50,
    extendedprice = k/100.0 (the doubles a 2-decimal text parses to), discount in {0.00..0.10}."""
    rng = np.random.default_rng(seed)
    qty = rng.integers(1, 51, n, dtype=np.int64)
    price = rng.integers(90000, 10500001, n, dtype=np.int64) / 100.0
    disc = rng.integers(0, 11, n, dtype=np.int64) / 100.0
    days = _ship_dates()
    ship = days[rng.integers(0, len(days), n)]
    return [Column(T_I64, qty), Column(T_F64, price), Column(T_F64, disc), Column(T_I64, ship)]


# ---- C1 ------------------------------------------------------------------------------------------------------
def c1_program(first_op_id: int = 100001) -> Program:
    sc = StageCompiler([T_I64], [None])
    sc.add_map(lambda x: x * x, first_op_id)
    sc.add_filter(lambda x: x % 2 == 0, first_op_id + 1)
    return sc.finish_memory()


# ---- C5 ------------------------------------------------------------------------------------------------------
def gen_keyed(n: int, n_keys: int, seed: int = 42, zipf: bool = False) -> List[Column]:
    """(key:str 'k%07d', v:i64) rows over n_keys distinct ids."""
    rng = np.random.default_rng(seed)
    if zipf:
        ids = np.minimum(rng.zipf(1.2, n) - 1, n_keys - 1).astype(np.int64)
    else:
        ids = rng.integers(0, n_keys, n, dtype=np.int64)
    digits = np.zeros((n, 8), dtype=np.uint8)
    digits[:, 0] = ord("k")
    x = ids.copy()
    for p in range(7, 0, -1):
        digits[:, p] = ord("0") + (x % 10)
        x //= 10
    offs = (np.arange(n + 1, dtype=np.int64) * 8).astype(np.uint32)
    vals = rng.integers(-1000, 1000, n, dtype=np.int64)
    return [Column(T_STR, digits.reshape(-1), offs), Column(T_I64, vals)]


def keyed_program(first_op_id: int = 100001) -> Program:
    sc = StageCompiler([T_STR, T_I64], ["key", "v"])
    return sc.finish_hash(["key"], lambda a, x: a + x[1], lambda a, b: a + b, 0, first_op_id)
