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
tdist.init_comm(local)
assert backend.comm_info(local) == (rank, world)
part_bits = st.run_host(local, mine, hi - lo).aggregate_bits()
total = ir.bits_f64(st.agg_finish(local, part_bits)[0])      # NCCL all-gather + rank-order fold inside the C ABI
(total_t,) = tdist.combine_aggregate([ir.bits_f64(part_bits[0])], [a.kind for a in prog.accs])
assert total == total_t, (total, total_t)                     # == the torch.distributed restatement of the same fold
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
tdist.exchange_hash_tables(hs, local)                          # owner-partitioned all-to-all on the device
fin = hs.hash_finish(local)
mine_k, mine_v = fin.column(0).to_values(), fin.column(1).to_values()
assert len(set(mine_k)) == len(mine_k)
# every rank holds the groups it owns: gather the shares (test plumbing) and compare the union with the oracle
import pickle
blob = pickle.dumps((mine_k, mine_v))
shares = [None] * world
dist.all_gather_object(shares, blob)
got = {}
n_total = 0
for b in shares:
    k, v = pickle.loads(b)
    n_total += len(k)
    got.update(zip(k, v))
ora = pyoracle.run_program(hp, kcols, m)
assert n_total == len(got), "a key is owned by two ranks"
assert got == dict(zip(ora.values(0), ora.values(1))), "rank %d: union of the owned groups differs from the oracle" % rank
assert 0 < len(mine_k) < len(got)
# a second exchange round on a fresh table: uniform keys, f64 partials are not used here (i64 sums are exact)
hs.hash_reset(local)
hs.run_host(local, [c.slice(lo, hi) for c in kcols], hi - lo).info
tdist.exchange_hash_tables(hs, local)
fin2 = hs.hash_finish(local)
assert dict(zip(fin2.column(0).to_values(), fin2.column(1).to_values())) == dict(zip(mine_k, mine_v))
dist.barrier()
backend.comm_destroy(local)
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


def test_two_gpus_one_process_local_comm(gpu):
    """One process driving two devices (tplx_gpu_comm_init_local = ncclCommInitAll), one host thread per device: the
    collectives of the C ABI (agg_finish, hash_exchange) against the oracle."""
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    from tuplex_b200 import backend, dist as tdist, ir, workloads as W
    from oracle import pyoracle
    if backend.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    devs = [0, 1]
    backend.init(devs)
    backend.comm_init_local(devs)
    try:
        n = 2_000_000
        cols = W.gen_lineitem(n, seed=9)
        prog = W.q6_program()
        st = backend.Stage(prog)

        def q6(r):
            lo, hi = tdist.shard_range(n, r, 2)
            bits = st.run_host(devs[r], [c.slice(lo, hi) for c in cols], hi - lo).aggregate_bits()
            return st.agg_finish(devs[r], bits)[0]
        with ThreadPoolExecutor(2) as ex:
            a, b = list(ex.map(q6, range(2)))
        assert a == b
        exp = None
        for r in range(2):
            lo, hi = tdist.shard_range(n, r, 2)
            v = ir.bits_f64(pyoracle.run_program(prog, [c.slice(lo, hi) for c in cols], hi - lo).acc_tree[0])
            exp = v if exp is None else exp + v
        assert ir.bits_f64(a) == exp
        m = 300_000
        kcols = W.gen_keyed(m, 5000, seed=3)
        hp = W.keyed_program()
        hs = backend.Stage(hp)

        def byk(r):
            lo, hi = tdist.shard_range(m, r, 2)
            hs.run_host(devs[r], [c.slice(lo, hi) for c in kcols], hi - lo).info
            hs.hash_exchange(devs[r])
            fin = hs.hash_finish(devs[r])
            return fin.column(0).to_values(), fin.column(1).to_values()
        with ThreadPoolExecutor(2) as ex:
            shares = list(ex.map(byk, range(2)))
        got = {}
        for k, v in shares:
            got.update(zip(k, v))
        ora = pyoracle.run_program(hp, kcols, m)
        assert sum(len(k) for k, _ in shares) == len(got)
        assert got == dict(zip(ora.values(0), ora.values(1)))
    finally:
        for dv in devs:
            backend.comm_destroy(dv)


def test_context_over_two_devices(gpu):
    """Context(tuplex.gpu.devices='0,1'): blocks sharded over both GPUs inside one process, outputs concatenated in order, the
    aggregate combined by tplx_gpu_agg_finish and aggregateByKey by tplx_gpu_stage_hash_exchange over the context's communicator."""
    import random
    import tuplex_b200
    from tuplex_b200 import backend
    if backend.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = random.Random(8)
    data = [(rng.randint(-50, 50), rng.randint(0, 6), "w%d" % rng.randint(0, 99)) for _ in range(60_000)]
    one = tuplex_b200.Context({"tuplex.gpu.blockRows": 7000})
    two = tuplex_b200.Context({"tuplex.gpu.blockRows": 7000, "tuplex.gpu.devices": "0,1"})
    pipe = lambda c: (c.parallelize(data, columns=["a", "b", "s"]).withColumn("q", lambda x: x["a"] // x["b"])
                       .resolve(ZeroDivisionError, lambda x: -999).filter(lambda x: x["q"] % 5 != 1)
                       .withColumn("t", lambda x: x["s"].upper() + str(x["q"])))
    assert pipe(two).collect() == pipe(one).collect()
    agg = lambda c: c.parallelize(data, columns=["a", "b", "s"]).aggregate(lambda x, y: x + y, lambda acc, r: acc + r["a"] * r["b"], 0).collect()
    assert agg(two) == agg(one) == [sum(a * b for a, b, _ in data)]
    fagg = lambda c: c.parallelize(data, columns=["a", "b", "s"]).aggregate(lambda x, y: x + y, lambda acc, r: acc + r["a"] * 0.1, 0.0).collect()
    assert abs(fagg(two)[0] - fagg(one)[0]) <= 1e-9 * abs(fagg(one)[0])
    byk = lambda c: c.parallelize(data, columns=["a", "b", "s"]).aggregateByKey(lambda x, y: x + y, lambda acc, r: acc + r["a"], 0, ["s"]).collect()
    assert sorted(byk(two)) == sorted(byk(one)) and len(byk(one)) == 100
    assert sorted(two.parallelize([r[2] for r in data]).unique().collect()) == sorted({r[2] for r in data})


def test_context_join_over_two_devices(gpu):
    """K8 with tuplex.gpu.devices='0,1': the build side is broadcast (one table per device), the probe blocks are sharded contiguously,
    the concatenation in device order is the probe order — equal to the one-device result."""
    import random
    import tuplex_b200
    from tuplex_b200 import backend
    if backend.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = random.Random(9)
    data = [(rng.randint(-50, 50), rng.randint(0, 6), "w%d" % rng.randint(0, 99)) for _ in range(60_000)]
    dim = [("w%d" % k, k * 10, None if k % 7 == 0 else float(k)) for k in range(0, 120, 2)]
    one = tuplex_b200.Context({"tuplex.gpu.blockRows": 7000})
    two = tuplex_b200.Context({"tuplex.gpu.blockRows": 7000, "tuplex.gpu.devices": "0,1"})

    def run(c):
        left = c.parallelize(data, columns=["a", "b", "s"])
        right = c.parallelize(dim, columns=["w", "v", "f"])
        return left.leftJoin(right, "s", "w", prefixes=(None, "d_")).collect()
    got1, got2 = run(one), run(two)
    assert got2 == got1 and len(got1) == len(data)
    assert sum(1 for r in got1 if r[3] is None) == sum(1 for r in data if int(r[2][1:]) % 2 == 1 or int(r[2][1:]) >= 120)


def test_cpp_host_two_devices(gpu, tmp_path):
    """C++ GpuBackend over two GPUs (--devices 0,1): tasks on both devices, tplx_gpu_agg_finish / tplx_gpu_stage_hash_exchange through
    the backend's NCCL communicator."""
    import struct
    import numpy as np
    from tuplex_b200 import backend, frontend, ir, workloads
    from tuplex_b200.backend import Column
    from tuplex_b200.ir import T_I64, T_STR
    from oracle import pyoracle
    from test_gpu_cpp_host import _decode_partition, _run_host
    if backend.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    lcols = workloads.load_lineitem_fixture()
    ln = len(lcols[0].data)
    qp = workloads.q6_program()
    info, prefix, _ = _run_host(tmp_path, qp, lcols, ln, [0, 1, 1, 0], 256 << 10, "0,1")
    (bits,) = struct.unpack("<q", open(prefix + ".agg", "rb").read())
    assert info["tasks"] == 2 and abs(ir.bits_f64(bits) - 1193053.2252999984) <= 1e-4
    rows = [("k%03d" % (i % 300), i % 7 - 3) for i in range(90_000)]
    kcols = [Column.from_values([r[0] for r in rows], T_STR), Column(T_I64, np.array([r[1] for r in rows], dtype=np.int64))]
    hp = frontend.StageCompiler([T_STR, T_I64], ["k", "v"]).finish_hash(["k"], lambda a, x: a + x[1], lambda a, b: a + b, 0, 100001)
    info, prefix, _ = _run_host(tmp_path, hp, kcols, len(rows), [T_STR, T_I64], 64 << 10, "0,1")
    got = []
    for i in range(info["hash_partitions"]):
        got += _decode_partition(open(f"{prefix}.hash{i}", "rb").read(), [T_STR, T_I64])
    want = {}
    for k, v in rows:
        want[k] = want.get(k, 0) + v
    assert dict(got) == want and len(got) == 300
