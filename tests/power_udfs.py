"""`**` with a literal integer exponent (reference: powerInst -> generateConstantIntegerPower, BlockGeneratorVisitor.cc:1313,5837)."""
import numpy as np

from tuplex_b200.backend import Column
from tuplex_b200.ir import T_F64, T_I64


def make_columns(n, seed):
    rnd = np.random.default_rng(seed)
    a = rnd.integers(-40, 41, n).astype(np.int64)
    a[::17] = 0
    f = np.round(rnd.normal(0, 30, n), 3)
    f[::13] = 0.0
    f[5::29] = -0.0
    return [Column(T_I64, a), Column(T_F64, f.astype(np.float64))], list(zip(a.tolist(), f.tolist()))


INT_CASES = [("lambda x: x['a'] ** %d" % k, k) for k in range(0, 7)]
NEG_CASES = [("lambda x: x['a'] ** %d" % k, k) for k in (-1, -2, -3)]
MIXED = ["lambda x: (x['a'] ** 2 + x['a'] ** 3) % 7", "lambda x: float(x['a'] ** 2) if x['a'] > 0 else x['a'] ** -1",
         "lambda x: (x['a'] > 3) ** 2 + 2 ** 3", "lambda x: 3 ** 2 * x['a']"]
FLOAT_CASES = [("lambda x: x['f'] ** %d" % k, k) for k in (0, 1, 2, 3, 4, 5, 6, -1, -2)]
UNSUPPORTED = ["lambda x: x['a'] ** 7", "lambda x: x['a'] ** x['a']", "lambda x: x['f'] ** 0.5", "lambda x: 2 ** x['a']"]
