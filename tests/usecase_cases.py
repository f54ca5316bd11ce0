"""Known-answer vectors of the reference's tuplex/test/core/UseCaseFunctionsTest.cc (line numbers beside each case): input rows, the
UDF source text exactly as the gtest passes it to UDF(...), and the rows the gtest expects. Shared by the CPU test (front end + oracle)
and the GPU test (through Context)."""
INF, NAN = float("inf"), float("nan")

extractOffer = ("def extractOffer(x):\n"
                "    offer = x.lower()\n"
                "\n"
                "    if 'sale' in offer:\n"
                "        offer = 'sale'\n"
                "    elif 'rent' in offer:\n"
                "        offer = 'rent'\n"
                "    elif 'sold' in offer:\n"
                "        offer = 'sold'\n"
                "    elif 'foreclos' in offer.lower():\n"
                "        offer = 'foreclosed'\n"
                "    else:\n"
                "        offer = 'unknown'\n"
                "\n"
                "    return offer")
extractPrice = ("def extractPrice(price, offer, facts, sqft):\n"
                "    p = 0\n"
                "    if offer == 'sold':\n"
                "        # price is to be calculated using price/sqft * sqft\n"
                "        val = facts\n"
                "        s = val[val.find('Price/sqft:') + len('Price/sqft:') + 1:]\n"
                "        r = s[s.find('$')+1:s.find(', ') - 1]\n"
                "        price_per_sqft = int(r)\n"
                "        p = price_per_sqft * sqft\n"
                "    elif offer == 'rent':\n"
                "        max_idx = price.rfind('/')\n"
                "        p = int(price[1:max_idx].replace(',', ''))\n"
                "    else:\n"
                "        # take price from price column\n"
                "        p = int(price[1:].replace(',', ''))\n"
                "\n"
                "    return p")
STRIP_ROWS = ["  hello ! \n ", "  \t\n", "abcde  \r\r", "  \n\r abcde"]
SEVEN = [3, 2, 1, 0, -1, -2, -3]

# (name, gtest line, rows, udf source, expected rows)
CASES = [
    ("LenCall", 38, ["hello", "world", "!", ""], "lambda x: len(x)", [5, 5, 1, 0]),
    ("UpperCall", 54, ["hello", "world", "!", ""], "lambda x: x.upper()", ["HELLO", "WORLD", "!", ""]),
    ("LowerCall", 68, ["HeLlo", "wOrld", "!", ""], "lambda x: x.lower()", ["hello", "world", "!", ""]),
    ("cleanCity", 82, ["WOBURN", "Woburn", "WINCHESTER", "Winchester", "SAUGUS", "Saugus", "Lynn", "DEDHAM", "Dedham", "BOSTON", "Boston", "Street"],
     "lambda x: x[0].upper() + x[1:].lower()",
     ["Woburn", "Woburn", "Winchester", "Winchester", "Saugus", "Saugus", "Lynn", "Dedham", "Dedham", "Boston", "Boston", "Street"]),
    ("IntCast", 99, ["200", "0", "-10", "42"], "lambda x: int(x)", [200, 0, -10, 42]),
    ("IntCastII", 123, [20.7, 3.141, 0.0, -8.7], "lambda x: int(x)", [20, 3, 0, -8]),
    ("FloatCast", 139, ["20", "3.141", "0", "iNf", "iNfINity", "naN", "-5.23"], "lambda x: float(x)", [20.0, 3.141, 0.0, INF, INF, NAN, -5.23]),
    ("FloatCast.int", 158, [20, -30], "lambda x: float(x)", [20.0, -30.0]),
    ("BoolCast.str", 164, ["hello", "3.141", "False", ""], "lambda x: bool(x)", [True, True, True, False]),
    ("BoolCast.int", 175, [20, -30, 0], "lambda x: bool(x)", [True, True, False]),
    ("BoolCast.float", 183, [-10.123, 0.0, 1.234, INF, NAN], "lambda x: bool(x)", [True, False, True, True, True]),
    ("StrCast.bool", 199, [(False, True)], "lambda x, y: (str(x), str(y))", [("False", "True")]),
    ("StrCast.float", 213, [-10.123, 3.141, 0.0, INF, NAN], "lambda x: str(x)", ["-10.123", "3.141", "0.0", "inf", "nan"]),
    ("StrCast.int", 226, [20, -30, 0], "lambda x: str(x)", ["20", "-30", "0"]),
    ("StringFormatOperator", 237, [12, 13, 14], "lambda x: '%04d' % x", ["0012", "0013", "0014"]),
    ("StringInOperator", 253, ["hello world", "what a wonderful world", "the earth is a globe"], "lambda x: 'world' in x", [True, True, False]),
    ("StringInOperatorII", 268, ["hello world", "what a wonderful world", "the earth is a globe"], "lambda x: 'world' not in x", [False, False, True]),
    ("extractOffer", 280, ["House for sale", "Townhouse for sale", "Condo for sale", "For sale by owner", "Apartment for sale", "Foreclosure",
                           "Foreclosed", "Coming soon", "New construction", "Make me move®"], extractOffer,
     ["sale"] * 5 + ["foreclosed"] * 2 + ["unknown"] * 3),
    ("strFindFunction", 325, [("hello", "l"), ("hello", "w"), ("hello", "")], "def test(x, y):\n    return x.find(y)", [2, -1, 0]),
    ("strReverseFindFunction", 347, [("/usr/local/hello", "/"), ("this.file.ext", "."), ("test", ""), ("", ""), ("/usr/local/hello", "\\")],
     "def test(x, y):\n    return x.rfind(y)", [10, 9, 4, 0, -1]),
    ("strReplaceFunction", 373, [("/usr/local/hello", "/"), ("hello world", "world"), ("test", ""), ("this is a test", "test"), ("hello world", "test")],
     "def test(x, y):\n    return x.replace(y, 'abc')", ["abcusrabclocalabchello", "hello abc", "abctabceabcsabctabc", "this is a abc", "hello world"]),
    ("strFormatFunction.v5", 462, [False, True], "lambda b: '{}'.format(b)", ["False", "True"]),
    ("strStrip", 492, STRIP_ROWS, "def test(s):\n    return s.strip()", ["hello !", "", "abcde", "abcde"]),
    ("strRstrip", 508, STRIP_ROWS, "lambda s: s.rstrip()", ["  hello !", "", "abcde", "  \n\r abcde"]),
    ("strLstrip", 520, STRIP_ROWS, "lambda s: s.lstrip()", ["hello ! \n ", "", "abcde  \r\r", "abcde"]),
    ("strStripChars", 532, [("!!abcdegt", "!atg"), ("www.test.com", "w."), ("hello", "helo"), ("?23test\n", "\nt?")], "lambda x, y: x.strip(y)",
     ["bcde", "test.com", "", "23tes"]),
    ("strSliceStartEnd", 577, [("hello", 0, 2), ("hello world", 2, 5)], "def test(x, a,b):\n    return x[a:b]", ["he", "llo"]),
    ("extractPrice", 599, [("$489,000", "sale", "2 bds , 1 ba , 920 sqft", 920), ("$3,250/mo", "rent", "3 bds , 1.5 ba , 1,100 sqft", 1100),
                           ("SOLD", "sold", "Price/sqft: $244 , 4 bds , 2 ba , 2,124 sqft", 2124)], extractPrice, [489000, 3250, 244 * 2124]),
    ("VariableOverwrite", 636, [10, 20], "def f(x):\n    x = 'hello'\n    return x", ["hello", "hello"]),
    ("VariableOverwriteIf", 655, [10, 2], "def f(x):\n   if x >= 10:\n       x = 'two digits'\n   else:\n       x = 'one digit'\n   return x",
     ["two digits", "one digit"]),
    ("IfShortCircuit.0", 735, SEVEN, "def g(x):\n    if True or (1/x > 0):\n        return 1\n    else:\n        return 0", [1] * 7),
    ("IfShortCircuit.1", 745, SEVEN, "def g(x):\n    if False and (1/x > 0):\n        return 1\n    else:\n        return 0", [0] * 7),
    ("IfShortCircuit.2", 755, SEVEN, "def g(x):\n    if (x == 1 or 1/(x-1) > 0) and (1/(x) > 0):\n        return 1\n    else:\n        return 0",
     [1, 1, 1, 0, 0, 0, 0]),
    ("IfShortCircuit.3", 766, SEVEN, "def l3(x):\n    if (x == 0 or x == 1) and ((x != 0 and 1/(x) <= 0) or (x != 1 and 1/(x-1) <= 0)):\n"
                                     "        return 1\n    else:\n        return 0", [0, 0, 0, 1, 0, 0, 0]),
]
# cases over named columns / other operators
COLUMN_CASES = [
    ("ColumnNamesMap", 682, [("hello", 20, -1, 9.0), ("world", 30, -2, 10.0), ("@", 40, -3, 11.0)], ["a", "b", "c", "d"], "map", None,
     "lambda x: (x['d'], x['b'])", [(9.0, 20), (10.0, 30), (11.0, 40)]),
    ("withColumnSimple", 696, [20, 30, 40], ["a"], "withColumn", "x", "lambda a: a / 2 - 1.0", [(20, 9.0), (30, 14.0), (40, 19.0)]),
    ("scientificNumbers", 713, [20000, 300, 4000, 100000], None, "filter", None, "lambda x: x < 2e4", [300, 4000]),
]
# Option inputs: NestedIf (:780, no golden in the gtest — CPython is the oracle) and FloatNullError (:950)
OPTION_CASES = [
    ("NestedIf", 780, [(89.0, None, None), (None, 1.0, None)], ["ActualElapsedTime", "DivReachedDest", "DivActualElapsedTime"], "withColumn", "ActualElapsedTime",
     "def fillInTimesUDF(row):\n    ACTUAL_ELAPSED_TIME = row['ActualElapsedTime']\n    if row['DivReachedDest']:\n        if int(row['DivReachedDest']) > 0:\n"
     "            return float(row['DivActualElapsedTime'])\n        else:\n            return ACTUAL_ELAPSED_TIME\n    else:\n        return ACTUAL_ELAPSED_TIME"),
    ("FloatNullError", 950, [None, None], None, "map", None, "lambda x: float(x) if x else None"),
]
