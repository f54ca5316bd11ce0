"""'Split at a marker' UDFs in the statement form the reference's benchmarks use (benchmarks/zillow/Z1/runtuplex.py:12-60:
find -> `if idx < 0: idx = len(s)` -> slice -> rfind -> `if idx < 0: 0 else idx + k` -> slice -> int). The front end fuses the two
idioms into TPLX_OP_SFINDE / TPLX_OP_SRFINDK (frontend.StageCompiler._fuse); the near misses below must NOT fuse.
Shared by the CPU test (front end + oracle vs CPython) and the GPU test (CUDA VM vs oracle)."""
import numpy as np

from tuplex_b200.backend import Column
from tuplex_b200.ir import T_I64, T_STR

CORPUS = ["3 bds , 2 ba , 1,560 sqft", "", "bd", " bd", "x bd, y bd", "1 bd", ",", ", ", "a, 12 bd, b", "12", "no marker here", "7 ba , 9 bd",
          "Studio , 1 ba , 500 sqft", "-- , 2 bds", "4 bds , -- ba", ",,,, 5 bd", "bd bd bd", "11 bds , 3.5 ba , 2,100 sqft", " ", "9", "a,b",
          "ends with comma,", ",starts", "ba ,", "2 ba , 880 sqft"]


def make_columns(n, seed):
    rnd = np.random.default_rng(seed)
    s = [CORPUS[i] for i in rnd.integers(0, len(CORPUS), n)]
    k = rnd.integers(-3, 4, n).astype(np.int64)
    return [Column.from_values(s, T_STR), Column(T_I64, k)], list(zip(s, k.tolist()))


def head_len(x):  # SFINDE
    s = x['s']
    stop = s.find(' bd')
    if stop < 0:
        stop = len(s)
    return stop


def number_before_marker(x):  # SFINDE + SRFINDK + int()
    s = x['s']
    stop = s.find(' bd')
    if stop < 0:
        stop = len(s)
    head = s[:stop]
    start = head.rfind(',')
    if start < 0:
        start = 0
    else:
        start += 2
    return int(head[start:])


def after_last_sep(x):  # SRFINDK with a longer needle and another K
    s = x['s']
    start = s.rfind('ba ,')
    if start < 0:
        start = 0
    else:
        start += 5
    return s[start:]


def both_as_expr(x):  # conditional-expression spelling of the same idioms
    s = x['s']
    i = s.find(',')
    j = s.rfind(' ')
    return (len(s) if i < 0 else i) * 100 + (0 if j < 0 else j + 1)


# near misses: same shape, different constants / operands -> must stay primitive sequences
def miss_default_one(x):
    s = x['s']
    start = s.rfind(',')
    if start < 0:
        start = 1
    else:
        start += 2
    return start


def miss_other_len(x):
    s = x['s']
    stop = s.find(' bd')
    if stop < 0:
        stop = len(s) - 1
    return stop


def miss_le(x):
    s = x['s']
    stop = s.find(',')
    if stop <= 0:
        stop = len(s)
    return stop


def miss_reuse(x):  # the raw find result is used again: the SFIND has to stay
    s = x['s']
    stop = s.find(',')
    raw = stop
    if stop < 0:
        stop = len(s)
    return stop * 1000 + raw


def miss_dynamic_k(x):
    s = x['s']
    start = s.rfind(',')
    if start < 0:
        start = 0
    else:
        start += x['k']
    return start


FUSED = [(head_len, {"SFINDE": 1}), (number_before_marker, {"SFINDE": 1, "SRFINDK": 1}), (after_last_sep, {"SRFINDK": 1}),
         (both_as_expr, {"SFINDE": 1, "SRFINDK": 1})]
MISSES = [miss_default_one, miss_other_len, miss_le, miss_reuse, miss_dynamic_k]
