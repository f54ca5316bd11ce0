"""BASELINE config 4 in miniature: the reference's flights pipeline (benchmarks/flights/runtuplex.py:131-260 — string splitting,
None-producing UDFs, an inner join with the carrier table, two left joins with the airport table, post-join cleaning, a filter on an
Option[int]) over synthetic tables with None values and rows that raise, through `Context` on the GPU, against the same pipeline in
plain CPython (the dual-mode invariant: results are Python's)."""
import numpy as np
import pytest

from tuplex_b200.pyexec import Row

from flights_udfs import cleanCode, divertedUDF, extractDefunctYear, fillInTimesUDF, filterDefunctFlights

F_COLS = ["Year", "OpUniqueCarrier", "Origin", "Dest", "OriginCityName", "DestCityName", "CrsArrTime", "CancellationCode", "Diverted",
          "Distance", "ActualElapsedTime", "DivReachedDest", "DivActualElapsedTime"]
C_COLS = ["Code", "Description"]
A_COLS = ["IATACode", "AirportName", "AirportCity", "LatitudeDecimal"]


def _tables(n, seed):
    rng = np.random.default_rng(seed)
    carriers = [("AA", "American Airlines Inc. (1960 - )"), ("PA", "Pan Am LLC (1927 - 1991)"), ("TW", "Trans World Co. (1930 - 2001)"),
                ("DL", "Delta Air Lines Inc. (1960 - )"), ("XX", "Broken (abc - )"), ("UA", "United Air Lines Inc. (1960 - )")]
    airports = [("ATL", "HARTSFIELD", "ATLANTA", 33.64), ("BOS", "LOGAN INTL", "BOSTON", 42.36), ("JFK", "JOHN F KENNEDY", "NEW YORK", 40.64),
                ("SFO", "SAN FRANCISCO INTL", "SAN FRANCISCO", 37.62), ("ORD", "OHARE", "CHICAGO", 41.98)]
    codes = ["AA", "PA", "TW", "DL", "XX", "UA", "ZZ"]   # ZZ: no carrier row -> dropped by the inner join
    ap = ["ATL", "BOS", "JFK", "SFO", "ORD", "LAX", "SEA"]  # LAX / SEA: no airport row -> None after the left joins
    city = {"ATL": "Atlanta, GA", "BOS": "Boston, MA", "JFK": "New York, NY", "SFO": "San Francisco, CA", "ORD": "Chicago, IL",
            "LAX": "Los Angeles, CA", "SEA": "Seattle, WA"}
    flights = []
    for i in range(n):
        o, d = ap[int(rng.integers(0, 7))], ap[int(rng.integers(0, 7))]
        div = float(rng.integers(0, 2)) if rng.random() < 0.1 else None
        flights.append((int(rng.integers(1987, 2020)), codes[int(rng.integers(0, 7))], o, d, city[o], city[d], int(rng.integers(0, 2400)),
                        [None, "A", "B", "C", "D", "E"][int(rng.integers(0, 6))] if rng.random() < 0.3 else None,
                        float(rng.integers(0, 2)), float(rng.integers(50, 3000)), float(rng.integers(30, 600)), div,
                        float(rng.integers(30, 700)) if div is not None else None))
    return flights, carriers, airports


def _apply(rows, names, kind, name, fn, ignore=()):
    """One operator on python rows, Tuplex semantics without resolvers: a row whose UDF raises is dropped."""
    out = []
    for r in rows:
        try:
            if kind == "withColumn":
                v = fn(Row(r, names))
                out.append(tuple(v if n == name else x for n, x in zip(names, r)) if name in names else r + (v,))
            elif kind == "mapColumn":
                i = names.index(name)
                out.append(r[:i] + (fn(r[i]),) + r[i + 1:])
            elif kind == "filter":
                if fn(Row(r, names)):
                    out.append(r)
        except Exception:  # noqa: BLE001
            pass
    return out, (names + [name] if kind == "withColumn" and name not in names else names)


def _join(L, ln, lk, R, rn, rk, left_outer, rprefix=""):
    li, ri = ln.index(lk), rn.index(rk)
    names = [n for n in ln if n != lk] + [lk] + [rprefix + n for n in rn if n != rk]
    out = []
    for l in L:
        ms = [r for r in R if r[ri] == l[li] and type(r[ri]) is type(l[li])]
        for r in ms:
            out.append(tuple(v for i, v in enumerate(l) if i != li) + (l[li],) + tuple(v for i, v in enumerate(r) if i != ri))
        if not ms and left_outer:
            out.append(tuple(v for i, v in enumerate(l) if i != li) + (l[li],) + (None,) * (len(rn) - 1))
    return out, names


def _expected(flights, carriers, airports):
    rows, names = list(flights), list(F_COLS)
    steps = [("withColumn", "OriginCity", lambda x: x['OriginCityName'][:x['OriginCityName'].rfind(',')].strip()),
             ("withColumn", "OriginState", lambda x: x['OriginCityName'][x['OriginCityName'].rfind(',') + 1:].strip()),
             ("withColumn", "DestCity", lambda x: x['DestCityName'][:x['DestCityName'].rfind(',')].strip()),
             ("mapColumn", "CrsArrTime", lambda x: '{:02}:{:02}'.format(int(x / 100), x % 100) if x else None),
             ("withColumn", "CancellationCode", cleanCode),
             ("mapColumn", "Diverted", lambda x: True if x > 0 else False),
             ("withColumn", "CancellationReason", divertedUDF),
             ("withColumn", "ActualElapsedTime", fillInTimesUDF)]
    for kind, name, fn in steps:
        rows, names = _apply(rows, names, kind, name, fn)
    crow, cn = list(carriers), list(C_COLS)
    for kind, name, fn in [("withColumn", "AirlineName", lambda x: x['Description'][:x['Description'].rfind('(')].strip()),
                           ("withColumn", "AirlineYearFounded", lambda x: int(x['Description'][x['Description'].rfind('(') + 1:x['Description'].rfind('-')])),
                           ("withColumn", "AirlineYearDefunct", extractDefunctYear)]:
        crow, cn = _apply(crow, cn, kind, name, fn)
    rows, names = _join(rows, names, "OpUniqueCarrier", crow, cn, "Code", False)
    rows, names = _join(rows, names, "Origin", list(airports), list(A_COLS), "IATACode", True, "Origin")
    rows, names = _join(rows, names, "Dest", list(airports), list(A_COLS), "IATACode", True, "Dest")
    rows, names = _apply(rows, names, "mapColumn", "Distance", lambda x: x / 0.00062137119224)
    rows, names = _apply(rows, names, "mapColumn", "AirlineName", lambda s: s.replace('Inc.', '').replace('LLC', '').replace('Co.', '').strip())
    names = ["CarrierName" if n == "AirlineName" else "OriginAirportIATACode" if n == "Origin" else n for n in names]
    rows, names = _apply(rows, names, "filter", None, filterDefunctFlights)
    return rows, names


@pytest.mark.gpu
def test_flights_like_pipeline_matches_cpython(gpu):
    import tuplex_b200 as tuplex
    flights, carriers, airports = _tables(4000, 3)
    want, want_names = _expected(flights, carriers, airports)
    assert 500 < len(want) < len(flights)
    c = tuplex.Context({"tuplex.gpu.blockRows": 1500})
    df = c.parallelize(flights, columns=F_COLS)
    df = df.withColumn('OriginCity', lambda x: x['OriginCityName'][:x['OriginCityName'].rfind(',')].strip())
    df = df.withColumn('OriginState', lambda x: x['OriginCityName'][x['OriginCityName'].rfind(',') + 1:].strip())
    df = df.withColumn('DestCity', lambda x: x['DestCityName'][:x['DestCityName'].rfind(',')].strip())
    df = df.mapColumn('CrsArrTime', lambda x: '{:02}:{:02}'.format(int(x / 100), x % 100) if x else None)
    df = df.withColumn('CancellationCode', cleanCode)
    df = df.mapColumn('Diverted', lambda x: True if x > 0 else False)
    df = df.withColumn('CancellationReason', divertedUDF)
    df = df.withColumn('ActualElapsedTime', fillInTimesUDF).ignore(TypeError)
    dc = c.parallelize(carriers, columns=C_COLS)
    dc = dc.withColumn('AirlineName', lambda x: x['Description'][:x['Description'].rfind('(')].strip())
    dc = dc.withColumn('AirlineYearFounded', lambda x: int(x['Description'][x['Description'].rfind('(') + 1:x['Description'].rfind('-')]))
    dc = dc.withColumn('AirlineYearDefunct', extractDefunctYear)
    da = c.parallelize(airports, columns=A_COLS)
    dall = df.join(dc, 'OpUniqueCarrier', 'Code')
    dall = dall.leftJoin(da, 'Origin', 'IATACode', prefixes=(None, 'Origin'))
    dall = dall.leftJoin(da, 'Dest', 'IATACode', prefixes=(None, 'Dest'))
    dall = dall.mapColumn('Distance', lambda x: x / 0.00062137119224)
    dall = dall.mapColumn('AirlineName', lambda s: s.replace('Inc.', '').replace('LLC', '').replace('Co.', '').strip())
    dall = dall.renameColumn('AirlineName', 'CarrierName').renameColumn('Origin', 'OriginAirportIATACode')
    dall = dall.filter(filterDefunctFlights)
    assert dall.columns == want_names
    got = dall.collect()
    assert len(got) == len(want)
    assert got == want
    assert c.metrics.kernel_launches > 0 and getattr(c.metrics, "join_probe_rows", 0) > 0  # the row stages and the joins ran on the device
