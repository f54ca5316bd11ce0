"""The C++ host mirror (tuplex_b200/host/gpu_backend.{h,cc}: IBackend::execute over the C ABI) end to end:
reference-format partitions in, reference-format output and exception partitions out, byte-identical to the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from tuplex_b200 import frontend
from tuplex_b200.backend import Column
from tuplex_b200.ir import T_I64, T_STR
from oracle import pyoracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_gpu_backend_execute(gpu, tmp_path):
    rng = np.random.default_rng(8)
    n = 40_000
    a = rng.integers(-30, 30, n, dtype=np.int64)
    s = ["id%d,%d" % (v, v * v) for v in a]
    cols = [Column(T_I64, a), Column.from_values(s, T_STR)]
    sc = frontend.StageCompiler([T_I64, T_STR], ["a", "s"])
    sc.add_with_column("q", lambda x: 1000 // x['a'], 100001)          # ZeroDivisionError rows
    sc.add_with_column("t", lambda x: x['s'][x['s'].find(',') + 1:].upper() + '#', 100002)
    sc.add_filter(lambda x: x['q'] % 3 != 1, 100003)
    prog = sc.finish_memory()
    psize = 256 << 10
    in_parts = pyoracle.to_partitions(cols, n, psize)   # what TransformStage::inputPartitions() would hold
    assert len(in_parts) > 3
    desc = tmp_path / "stage.bin"
    desc.write_bytes(prog.serialize())
    part_files = []
    for i, p in enumerate(in_parts):
        f = tmp_path / f"in{i}.bin"
        f.write_bytes(p)
        part_files.append(str(f))
    exe = os.path.join(ROOT, "tuplex_b200", "lib", "tplx_host_run")
    prefix = str(tmp_path / "res")
    r = subprocess.run([exe, str(desc), "0,3", str(psize), prefix] + part_files, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    ora = pyoracle.run_program(prog, cols, n)
    assert info["out_rows"] == ora.n_out and info["exceptions"] == len(ora.exceptions) and info["exceptions"] > 0

    class OC:
        def __init__(self, t, d, o): self.type, self.data, self.offsets = t, d, o
    want_parts = pyoracle.to_partitions([OC(*c) for c in ora.columns], ora.n_out, psize)
    assert info["out_partitions"] == len(want_parts)
    for i, w in enumerate(want_parts):
        assert open(f"{prefix}.out{i}", "rb").read() == w, f"output partition {i}"
    assert open(prefix + ".exc", "rb").read() == pyoracle.exception_partition(cols, ora.exceptions)
    # stage-level failure surfaces as an error, not a crash
    bad = tmp_path / "bad.bin"
    bad.write_bytes(prog.serialize()[:-16])
    r2 = subprocess.run([exe, str(bad), "0,3", str(psize), prefix] + part_files, capture_output=True, text=True, timeout=120)
    assert r2.returncode == 1 and "stage descriptor" in r2.stderr
