"""Test vectors restated from the reference's Python-level tests (tuplex/python/tests/*.py): input rows, UDF, expected
result of `Context.parallelize(rows).map(udf).collect()`. Where the reference test computes the expectation with CPython
(`list(map(g, rows))`) the same is done here. Each entry: (name + reference test, rows, udf, expected-or-None, columns).
expected None = evaluate `udf` with CPython on every row (multi-parameter lambdas get the row unpacked)."""

HELLO = "hello"
_idx = [-10, -2, 3, 1, 10]
_pairs = [(-10, 10), (-10, -2), (10, -2), (-10, 2), (2, -10), (-10, 3), (-2, 10), (-2, -1), (-3, 4), (1, 10), (1, -3), (1, 3)]
_x = [-10, -3, 0, 5, 7]
_y = [-7, 2, 3, 10]
_sh = [0, 1, 2, 3, 4]

MAP_CASES = [
    # ---- test_strings.py ------------------------------------------------------------------------------------------
    ("strings.concat :22", [("hello", "world"), ("foo", "bar"), ("blank", ""), ("", "another"), ("", "")], lambda a, b: a + b,
     ["helloworld", "foobar", "blank", "another", ""]),
    ("strings.slice s[i:] :74", [(HELLO, i) for i in _idx], lambda s, i: s[i:], None),
    ("strings.slice s[:i] :80", [(HELLO, i) for i in _idx], lambda s, i: s[:i], None),
    ("strings.slice a[b:y] :91", [(HELLO, a, b) for a, b in _pairs], lambda a, b, y: a[b:y], None),
    ("strings.strcast bools :115", [(False, True)], lambda x: (x[0], x[1]), [(False, True)]),
    ("strings.strcast str(int) :127", [-10, 0, 20], lambda x: str(x), ["-10", "0", "20"]),
    ("strings.strcast str(str) :129", ["-10", "hello", "", "   bye   ", "7.123"], lambda x: str(x), None),
    ("strings.strip :153", [" \n\t\r hello \n\r\r\r  ", "  \t\r\x0b", "goodbye!\n\n\n", " \r\tabcde"], lambda x: x.strip(), None),
    ("strings.lstrip :154", [" \n\t\r hello \n\r\r\r  ", "  \t\r\x0b", "goodbye!\n\n\n", " \r\tabcde"], lambda x: x.lstrip(), None),
    ("strings.rstrip :155", [" \n\t\r hello \n\r\r\r  ", "  \t\r\x0b", "goodbye!\n\n\n", " \r\tabcde"], lambda x: x.rstrip(), None),
    ("strings.strip twice :161", [" \n\t\r hello \n\r\r\r  ", "  \t\r\x0b", "goodbye!\n\n\n", " \r\tabcde"], lambda x: x.lstrip().rstrip(), None),
    ("strings.startswith :166", [("hello", "h"), ("hello", "he"), ("Hello", "hello"), ("abcde", "abcde")], lambda s, p: s.startswith(p),
     [True, True, False, True]),
    # ---- test_arithmetic.py ---------------------------------------------------------------------------------------
    ("arith.add int+float :26", [1, 2, 4], lambda x: x + 10.7, [11.7, 12.7, 14.7]),
    ("arith.unary plus :29", [1, 2, 4], lambda x: +x, [1, 2, 4]),
    ("arith.unary plus float :35", [1.1, 2.1, 4.1], lambda x: +x, [1.1, 2.1, 4.1]),
    ("arith.sub :40", [11, 12, 13], lambda x: x - 10, [1, 2, 3]),
    ("arith.neg :43", [11, 12, 13], lambda x: -x, [-11, -12, -13]),
    ("arith.sub bool :46", [0], lambda x: x - True, [-1]),
    ("arith.sub int-float :61", [1, 2, 3], lambda x: x - 0.1, [0.9, 1.9, 2.9]),
    ("arith.invert :69", [0, 1, 2], lambda x: ~x, [-1, -2, -3]),
    ("arith.div :77", [0, 1, 2, -5, -10], lambda x: x / 10, [0, 0.1, 0.2, -0.5, -1.0]),
    ("arith.div by zero rows dropped :81", [(0, 0), (-1, 0), (1, 0), (42, 1)], lambda x: x[0] / x[1], [42]),
    ("arith.idiv :86", [10, 11, 12, 13, 14, 15, 16], lambda x: x // 7, [1, 1, 1, 1, 2, 2, 2]),
    ("arith.idiv float :89", [10, 11, 12, 13, 14, 15, 16], lambda x: x // 7.0, [1.0, 1.0, 1.0, 1.0, 2.0, 2.0, 2.0]),
    ("arith.idiv neg :92", [-10, -9, -8, -7, -6, -5], lambda x: x // 6, [-2, -2, -2, -2, -1, -1]),
    ("arith.idiv neg neg :95", [-10, -9, -8, -7, -6, -5], lambda x: x // -6, [1, 1, 1, 1, 1, 0]),
    ("arith.idiv neg float :98", [-10, -9, -8, -7, -6, -5], lambda x: x // -6.0, [1.0, 1.0, 1.0, 1.0, 1.0, 0.0]),
    ("arith.mod i64 i64 :117", [(x, y) for x in _x for y in _y], lambda a, b: a % b, None),
    ("arith.mod i64 f64 :128", [(x, float(y)) for x in _x for y in _y], lambda a, b: a % b, None),
    ("arith.mod f64 i64 :139", [(float(x), y) for x in _x for y in _y], lambda a, b: a % b, None),
    ("arith.mod f64 f64 :150", [(float(x), float(y)) for x in _x for y in _y], lambda a, b: a % b, None),
    ("arith.lshift :165", [(x, y) for x in _x for y in _sh], lambda a, b: a << b, None),
    ("arith.rshift", [(x, y) for x in _x for y in _sh], lambda a, b: a >> b, None),
    # ---- test_filter.py / test_tuples.py / test_index.py style vectors -------------------------------------------------
    ("tuples.swap", [(1, "a"), (2, "b")], lambda x: (x[1], x[0]), [("a", 1), ("b", 2)]),
    ("tuples.nested build", [(1, 2), (3, 4)], lambda a, b: (a + b, a * b, a - b), [(3, 2, -1), (7, 12, -1)]),
    ("index.negative", [("abc", "xyz"), ("de", "uvw")], lambda x: x[-1][-1] + x[0][0], ["za", "wd"]),
    ("logical.and or not", [(True, False), (False, False), (True, True)], lambda a, b: (a and b, a or b, not a), None),
    ("logical.compare chain", [1, 5, 10, 15], lambda x: 2 < x <= 10, [False, True, True, False]),
    ("logical.ternary", [-2, 0, 3], lambda x: "neg" if x < 0 else ("zero" if x == 0 else "pos"), ["neg", "zero", "pos"]),
]

# (name, rows, columns, pipeline builder, expected)  — tuplex/python/tests/test_columns.py
COLUMN_CASES = [
    ("columns.withColumnNew :25", [10, 20, 3, 4], None, lambda ds: ds.withColumn("newcol", lambda x: 2 * x), [(10, 20), (20, 40), (3, 6), (4, 8)]),
    ("columns.withColumnSame :29", [(1, "Hello"), (2, "world")], ["count", "word"],
     lambda ds: ds.withColumn("word", lambda x: x["word"][-1] * x["count"]), [(1, "o"), (2, "dd")]),
    ("columns.withColumnSameII :36", [(1, "Hello"), (2, "world")], ["count", "word"],
     lambda ds: ds.withColumn("word", lambda x: x[1][-1] * x[0]), [(1, "o"), (2, "dd")]),
    ("columns.mapColumn :43", [1, 2, 3], ["A"], lambda ds: ds.mapColumn("A", lambda x: x + 1), [2, 3, 4]),
    ("columns.select two :50", [(1, 2, 3), (4, 5, 6), (7, 8, 9)], ["abc", "def", "ghi"], lambda ds: ds.selectColumns(["abc", "ghi"]), [(1, 3), (4, 6), (7, 9)]),
    ("columns.select one :52", [(1, 2, 3), (4, 5, 6), (7, 8, 9)], ["abc", "def", "ghi"], lambda ds: ds.selectColumns(["abc"]), [1, 4, 7]),
    ("columns.select index :66", [(1, 2, 3), (4, 5, 6), (7, 8, 9)], ["abc", "def", "ghi"], lambda ds: ds.selectColumns(2), [3, 6, 9]),
    ("columns.select neg index :68", [(1, 2, 3), (4, 5, 6), (7, 8, 9)], ["abc", "def", "ghi"], lambda ds: ds.selectColumns(-2), [2, 5, 8]),
    ("columns.select doubled :72", [(1, 2, 3), (4, 5, 6), (7, 8, 9)], ["abc", "def", "ghi"], lambda ds: ds.selectColumns(["abc", "abc"]), [(1, 1), (4, 4), (7, 7)]),
    ("columns.select ints :76", [(1, 2, 3), (4, 5, 6), (7, 8, 9)], ["abc", "def", "ghi"], lambda ds: ds.selectColumns([1, 0]), [(2, 1), (5, 4), (8, 7)]),
    ("columns.select mixed :80", [(1, 2, 3), (4, 5, 6), (7, 8, 9)], ["abc", "def", "ghi"], lambda ds: ds.selectColumns([-1, "def", "ghi"]),
     [(3, 2, 3), (6, 5, 6), (9, 8, 9)]),
    ("columns.withColumnUnnamed :84", [(1, 2), (3, 2)], None, lambda ds: ds.withColumn("newcol", lambda a, b: (a + b) / 10), [(1, 2, 3 / 10), (3, 2, 5 / 10)]),
]


def expected_of(rows, udf, expected):
    if expected is not None:
        return expected
    import inspect
    n = len(inspect.signature(udf).parameters)
    return [udf(*r) if (n > 1 and isinstance(r, tuple)) else udf(r) for r in rows]
