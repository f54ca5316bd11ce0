"""GPU fuzz: random UDF pipelines through the C ABI vs the oracle, bit-exact (rows, strings, exception records),
with and without the prefilter hint, at a size that spans many tiles."""
import numpy as np
import pytest

from tuplex_b200 import backend, frontend
from oracle import pyoracle
from fuzz_udfs import COLS, TYPES, Gen, apply_ops, make_columns
from helpers import assert_result_equals_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_random_pipelines_match_oracle(gpu, seed):
    g = Gen(1000 + seed)
    n = 20_000 + 37 * seed
    cols, _ = make_columns(n, seed)
    compared = 0
    for trial in range(10):
        ops = g.pipeline()
        progs = []
        for pre in (True, False):
            sc = frontend.StageCompiler(TYPES, COLS)
            try:
                apply_ops(sc, ops)
                progs.append(sc.finish_memory(prefilter=pre))
            except frontend.UnsupportedUDF:
                progs = []
                break
        if not progs:
            continue
        ora = pyoracle.run_program(progs[1], cols, n, first_row_no=seed)
        for prog in progs if progs[0].prefilter is not None else progs[1:]:
            st = backend.Stage(prog)
            res = st.run_host(0, cols, n, first_row_no=seed)
            assert_result_equals_oracle(res, ora, f"seed {seed} trial {trial} prefilter={prog.prefilter is not None}: {ops}")
            res.free()
            st.close()
        compared += 1
    assert compared >= 3
