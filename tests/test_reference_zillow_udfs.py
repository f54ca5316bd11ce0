"""The reference's LITERAL Z1 UDFs (tests/golden/zillow_z1_udfs_ref.py = benchmarks/zillow/Z1/runtuplex.py:12-105, copied verbatim by
tests/golden/make_golden.py) through this repo's front end: closes the common-mode hole of testing only the re-worded UDFs of
tuplex_b200/workloads.py. CPU: front end + oracle -> the reference's golden output (md5 of zillow.cpp / runpython.py).
GPU (marked): the same program on the device, bit-equal to the oracle and to the golden."""
import hashlib
import importlib.util
import os

import pytest

from tuplex_b200 import frontend, workloads
from tuplex_b200.ir import OP_NAMES
from oracle import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_MD5 = "4d5ca0263b1a5058341a369116dee83a"


def _ref_udfs():
    spec = importlib.util.spec_from_file_location("zillow_z1_udfs_ref", os.path.join(HERE, "golden", "zillow_z1_udfs_ref.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def reference_z1_program(first_op_id=100001):
    """The operator chain of benchmarks/zillow/Z1/runtuplex.py:192-205 with the reference's own UDF objects and lambdas."""
    R = _ref_udfs()
    sc = frontend.StageCompiler(workloads.ZILLOW_TYPES, workloads.ZILLOW_COLS)
    k = first_op_id
    sc.add_with_column("bedrooms", R.extractBd, k)
    sc.add_filter(lambda x: x['bedrooms'] < 10, k + 1)
    sc.add_with_column("type", R.extractType, k + 2)
    sc.add_filter(lambda x: x['type'] == 'house', k + 3)
    sc.add_with_column("zipcode", lambda x: '%05d' % int(x['postal_code']), k + 4)
    sc.add_map_column("city", lambda x: x[0].upper() + x[1:].lower(), k + 5)
    sc.add_with_column("bathrooms", R.extractBa, k + 6)
    sc.add_with_column("sqft", R.extractSqft, k + 7)
    sc.add_with_column("offer", R.extractOffer, k + 8)
    sc.add_with_column("price", R.extractPrice, k + 9)
    sc.add_filter(lambda x: 100000 < x['price'] < 2e7, k + 10)
    sc.add_select(["url", "zipcode", "address", "city", "state", "bedrooms", "bathrooms", "sqft", "offer", "type", "price"], k + 11)
    return sc.finish_memory()


def _md5_of(columns_values):
    return hashlib.md5(workloads.rows_to_csv(columns_values, workloads.ZILLOW_OUT)).hexdigest()


def test_literal_reference_udfs_through_frontend_and_oracle(built):
    prog = reference_z1_program()
    # the reference's statement form hits the fused idioms (find / `if idx < 0` / rfind / `+= 2`)
    names = [OP_NAMES[i.op] for i in prog.instrs]
    assert names.count("SFINDE") == 3 and names.count("SRFINDK") == 3, names
    cols, n = workloads.load_zillow_fixture()
    ora = pyoracle.run_program(prog, cols, n)
    assert ora.n_out == 577 and len(ora.exceptions) == 0
    assert _md5_of([ora.values(c) for c in range(len(ora.columns))]) == GOLDEN_MD5
    # and the separately declared filter UDFs of the script (filterBd / filterType / filterPrice with its `<=`) lower too
    R = _ref_udfs()
    sc = frontend.StageCompiler(workloads.ZILLOW_TYPES, workloads.ZILLOW_COLS)
    sc.add_with_column("bedrooms", R.extractBd, 1)
    sc.add_filter(R.filterBd, 2)
    sc.add_with_column("type", R.extractType, 3)
    sc.add_filter(R.filterType, 4)
    sc.add_with_column("bathrooms", R.extractBa, 5)
    sc.add_with_column("sqft", R.extractSqft, 6)
    sc.add_with_column("offer", R.extractOffer, 7)
    sc.add_with_column("price", R.extractPrice, 8)
    sc.add_filter(R.filterPrice, 9)
    sc.add_select(["bedrooms", "price"], 10)
    p2 = sc.finish_memory()
    o2 = pyoracle.run_program(p2, cols, n)
    assert o2.n_out == 577  # no listing of the fixture costs exactly 2e7


@pytest.mark.gpu
def test_literal_reference_udfs_on_the_device(gpu):
    from helpers import assert_result_equals_oracle, run_both
    prog = reference_z1_program()
    cols, n = workloads.load_zillow_fixture()
    st, res, ora = run_both(prog, cols, n)
    assert_result_equals_oracle(res, ora, "literal reference Z1 UDFs")
    assert _md5_of([c.to_values() for c in res.columns()]) == GOLDEN_MD5
