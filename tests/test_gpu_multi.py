"""Two ranks, one GPU each, NCCL: shard blocks, run the stage per rank, combine (aggregate) / exchange (aggregateByKey),
check against the single-process oracle. Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["TPLX_ROOT"])
import numpy as np, torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from tuplex_b200 import backend, dist as tdist, ir, workloads as W
from oracle import pyoracle
backend.init([local])
# ---- Q6: shard rows contiguously, combine partials in rank order --------------------------------------------
n = 3_000_000
cols = W.gen_lineitem(n, seed=5)
lo, hi = tdist.shard_range(n, rank, world)
mine = [c.slice(lo, hi) for c in cols]
prog = W.q6_program()
st = backend.Stage(prog)
part = ir.bits_f64(st.run_host(local, mine, hi - lo).aggregate_bits()[0])
(total,) = tdist.combine_aggregate([part], [a.kind for a in prog.accs])
# expected: per-shard oracle trees combined in rank order
exp = None
for r in range(world):
    a, b = tdist.shard_range(n, r, world)
    v = ir.bits_f64(pyoracle.run_program(prog, [c.slice(a, b) for c in cols], b - a).acc_tree[0])
    exp = v if exp is None else exp + v
assert total == exp, (total, exp)
seq = ir.bits_f64(pyoracle.run_program(prog, cols, n).acc_seq[0])
assert abs(total - seq) <= 1e-9 * abs(seq)
# ---- aggregateByKey: per-rank tables, all-gather + merge, every rank holds the global result ------------------
m = 400_000
kcols = W.gen_keyed(m, 3000, seed=11)
lo, hi = tdist.shard_range(m, rank, world)
hp = W.keyed_program()
hs = backend.Stage(hp)
hs.run_host(local, [c.slice(lo, hi) for c in kcols], hi - lo).info
tdist.exchange_hash_tables(hs, local)
fin = hs.hash_finish(local)
got = dict(zip(fin.column(0).to_values(), fin.column(1).to_values()))
ora = pyoracle.run_program(hp, kcols, m)
assert got == dict(zip(ora.values(0), ora.values(1))), "rank %d table differs" % rank
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", total)
'''


def test_two_gpus_nccl(gpu, tmp_path):
    from tuplex_b200 import backend
    if backend.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, TPLX_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
