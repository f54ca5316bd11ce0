"""Selective filter chains that the planner can state as a string-scan hint (tplx_scan_term, include/tplx_ir.h) and near misses that
it must not. Shared by the CPU test (matcher + oracle vs CPython) and the GPU test (K1f closed form == VM == oracle)."""
import numpy as np

from tuplex_b200.backend import Column
from tuplex_b200.ir import T_F64, T_I64, T_STR

FACTS = ["3 bds , 2 ba , 1,560 sqft", "", "bd", " bd", "x bd, y bd", "1 bd", ",", ", ", "a, 12 bd, b", "12", "no marker here", "7 ba , 9 bd",
         "Studio , 1 ba , 500 sqft", "-- , 2 bds", "4 bds , -- ba", ",,,, 5 bd", "11 bds , 3.5 ba", " ", "9", "a,b", ", 10 bds", "  8 bd", ", -3 bd"]
TITLES = ["House For Sale", "HOUSE", "house", "Condo for rent", "Apartment", "townhouse sold", "", "hous", "ouse", "Lot/Land", "New HoUsE!",
          "Foreclosed home", "houSe boat", "x" * 19 + "house", "ho" * 9]


def make_columns(n, seed):
    rnd = np.random.default_rng(seed)
    f = [FACTS[i] for i in rnd.integers(0, len(FACTS), n)]
    t = [TITLES[i] for i in rnd.integers(0, len(TITLES), n)]
    k = rnd.integers(-5, 15, n).astype(np.int64)
    g = np.round(rnd.normal(5, 4, n), 2)
    g[::41] = float("nan")
    return [Column.from_values(f, T_STR), Column.from_values(t, T_STR), Column(T_I64, k), Column(T_F64, g.astype(np.float64))]


TYPES = [T_STR, T_STR, T_I64, T_F64]
NAMES = ["facts", "title", "k", "g"]


def bedrooms(x):
    val = x['facts']
    max_idx = val.find(' bd')
    if max_idx < 0:
        max_idx = len(val)
    s = val[:max_idx]
    split_idx = s.rfind(',')
    if split_idx < 0:
        split_idx = 0
    else:
        split_idx += 2
    return int(s[split_idx:])


def baths_neg_skip(x):  # another marker / separator / skip (negative: Python's negative slice start)
    val = x['facts']
    i = val.find(' ba')
    if i < 0:
        i = len(val)
    s = val[:i]
    j = s.rfind(', ')
    if j < 0:
        j = 0
    else:
        j += -1
    return int(s[j:])


def heavy_tail(sc, k):
    """work behind the selective filter, so that the planner splits a prefilter off"""
    sc.add_with_column("z1", lambda x: x['title'].replace('o', '0') + '|' + x['facts'].upper(), k)
    sc.add_with_column("z2", lambda x: x['z1'].find('BD') + len(x['z1'].replace(',', '')), k + 1)
    sc.add_with_column("z3", lambda x: '%05d' % x['z2'] + x['title'].lower(), k + 2)


# (name, builder(sc) adding the selective head, expected term kinds) — kinds: 0 CONTAINS, 1 FIELD_INT, 2 FIXED
def _h_zillow(sc):
    sc.add_with_column("bedrooms", bedrooms, 100001)
    sc.add_filter(lambda x: x['bedrooms'] < 10, 100002)
    sc.add_filter(lambda x: 'house' in x['title'].lower(), 100003)


def _h_contains_only(sc):
    sc.add_filter(lambda x: 'HOUSE' in x['title'].upper(), 100001)


def _h_not_contains_raw(sc):
    sc.add_filter(lambda x: 'ou' not in x['title'], 100001)
    sc.add_filter(lambda x: ',' in x['facts'], 100002)


def _h_fixed_then_field(sc):
    sc.add_filter(lambda x: x['k'] >= 3, 100001)
    sc.add_filter(lambda x: 2.5 < x['g'], 100002)
    sc.add_with_column("b", baths_neg_skip, 100003)
    sc.add_filter(lambda x: 1 != x['b'], 100004)


def _h_field_eq_and_nan(sc):
    sc.add_with_column("bedrooms", bedrooms, 100001)
    sc.add_filter(lambda x: x['bedrooms'] == 3, 100002)
    sc.add_filter(lambda x: x['g'] != 5.0, 100003)   # FCMP_ONE: false for NaN (reference quirk)


HINTED = [("zillow", _h_zillow, [1, 0]), ("contains_upper", _h_contains_only, [0]), ("not_contains_raw", _h_not_contains_raw, [0, 0]),
          ("fixed_then_field", _h_fixed_then_field, [2, 2, 1]), ("field_eq_nan", _h_field_eq_and_nan, [1, 2])]


def _m_value_unused(sc):  # int() whose value no filter tests: can raise, the closed form would lose it
    sc.add_with_column("bedrooms", bedrooms, 100001)
    sc.add_filter(lambda x: 'house' in x['title'].lower(), 100002)


def _m_or(sc):
    sc.add_filter(lambda x: 'house' in x['title'].lower() or x['k'] > 12, 100001)


def _m_arith(sc):
    sc.add_with_column("bedrooms", bedrooms, 100001)
    sc.add_filter(lambda x: x['bedrooms'] * 2 < 10, 100002)


def _m_two_columns(sc):
    sc.add_filter(lambda x: x['k'] < x['g'], 100001)


MISSES = [("value_unused", _m_value_unused), ("or", _m_or), ("arith", _m_arith), ("two_columns", _m_two_columns)]
