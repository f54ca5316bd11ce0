"""world_size-2 gloo run of the one exchange step of the path (aggregate combine, table all-gather)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["TPLX_ROOT"])
import numpy as np
import torch.distributed as dist
from tuplex_b200 import dist as tdist, ir
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
kinds = [ir.C["TPLX_ACC_SUM_F64"], ir.C["TPLX_ACC_SUM_I64"], ir.C["TPLX_ACC_MAX_I64"]]
vals = [0.1 * (rank + 1), 10 ** 18 * (rank + 7), rank * 5 - 3]
out = tdist.combine_aggregate(vals, kinds)
exp_f = 0.0
for r in range(world):
    exp_f = (0.1 * (r + 1)) if r == 0 else exp_f + 0.1 * (r + 1)
exp_i = sum(10 ** 18 * (r + 7) for r in range(world))
exp_i = (exp_i + 2 ** 63) % 2 ** 64 - 2 ** 63
assert out[0] == exp_f and out[1] == exp_i and out[2] == (world - 1) * 5 - 3, out
# variable-length all-gather of (values, offsets, bytes)-style arrays
a = np.arange(rank * 3 + 1, dtype=np.int64) + 100 * rank
b = np.frombuffer(("k%d" % rank).encode() * (rank + 1), dtype=np.uint8)
g = tdist.allgather_arrays([a, b])
for r in range(world):
    assert np.array_equal(g[r][0], np.arange(r * 3 + 1, dtype=np.int64) + 100 * r)
    assert g[r][1].tobytes() == ("k%d" % r).encode() * (r + 1)
lo, hi = tdist.shard_range(10, rank, world)
assert (lo, hi) == ((0, 5) if rank == 0 else (5, 10))
dist.barrier()
dist.destroy_process_group()
os.write(1, ("rank %d ok\n" % rank).encode())  # one write: the two ranks share the pipe
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TPLX_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
    for attempt in range(2):  # the free port found below can be taken by another process before torchrun binds it
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", str(port), str(script)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
        if r.returncode == 0 or "address already in use" not in (r.stdout + r.stderr).lower():
            break
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
