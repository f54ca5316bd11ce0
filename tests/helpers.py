"""Shared comparison helpers for the parity tests (GPU result vs CPU oracle, bit-exact)."""
import numpy as np

from tuplex_b200 import backend, ir
from oracle import pyoracle


def assert_result_equals_oracle(res: backend.Result, ora: pyoracle.OracleResult, what=""):
    info = res.info
    assert int(info.n_out_rows) == ora.n_out, f"{what}: row count {info.n_out_rows} != oracle {ora.n_out}"
    assert int(info.n_exceptions) == len(ora.exceptions), f"{what}: exception count {info.n_exceptions} != {len(ora.exceptions)}"
    for c, (t, odata, ooffs) in enumerate(ora.columns):
        col = res.column(c)
        assert col.type == t, f"{what}: column {c} type"
        if t == ir.T_STR:
            assert np.array_equal(col.offsets, ooffs), f"{what}: column {c} string offsets differ"
            assert col.data.tobytes() == odata.tobytes(), f"{what}: column {c} string bytes differ"
        else:
            assert np.array_equal(col.data.view(np.int64), odata.view(np.int64)), f"{what}: column {c} values differ (bit compare)"
    exc = res.exceptions()
    for f in ("row", "row_no", "code", "op_id"):
        assert np.array_equal(exc[f], ora.exceptions[f]), f"{what}: exception field {f} differs"


def run_both(prog, cols, n, first_row_no=0, device=0):
    st = backend.Stage(prog)
    res = st.run_host(device, cols, n, first_row_no)
    ora = pyoracle.run_program(prog, cols, n, first_row_no)
    return st, res, ora
