"""Multi-GPU plumbing: one process per GPU, torch.distributed for the one exchange step of the path.

The reference's only intra-node strategy is data parallelism over partitions (one TransformTask per
partition on executor threads, tuplex/core/src/ee/local/LocalBackend.cc:679-735,1531-1586) followed by a
sequential combine of partial aggregates on the driver (TransformTask.cc:278-299; per-task hash tables merged
pairwise, LocalBackend.cc:2288-2375). Here: blocks are sharded contiguously over ranks (so that concatenating
per-rank outputs in rank order preserves input order, LocalBackend.cc:1104-1152), the map/filter phase needs
no communication, and only aggregate endpoints exchange data:
  aggregate       -> ncclAllGather of the per-rank partial + fold in rank order on the device (tplx_gpu_agg_finish)
  aggregateByKey  -> owner = hash(key) mod world, grouped ncclSend/ncclRecv all-to-all of packed (key, partial) records
                     between device buffers, local merge (tplx_gpu_stage_hash_exchange); nothing passes through the host
Both live inside the C ABI (csrc/tplx_gpu_comm.inl); this module only bootstraps the communicator (the 128-byte NCCL id travels
over the process group the launcher already set up) and shards work. combine_aggregate / allgather_arrays are the
torch.distributed restatement of the same rank-order fold, kept for the gloo (CPU) tests of the host-side logic.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .ir import C


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_comm(device: int, group=None):
    """Create this rank's NCCL communicator inside libtplx_gpu.so. Collective over the torch.distributed group: rank 0
    draws the id, every rank joins with its own device."""
    import torch
    import torch.distributed as dist
    from . import backend
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if backend.comm_info(device) is not None:
        return rank, world
    uid = backend.comm_unique_id() if rank == 0 else bytes(backend.COMM_ID_BYTES)
    t = torch.tensor(list(uid), dtype=torch.uint8, device=_device_for(dist))
    dist.broadcast(t, src=0, group=group)
    backend.comm_init(device, rank, world, bytes(t.cpu().tolist()))
    return rank, world


def _combine(kind: int, a, b):
    if kind == C["TPLX_ACC_SUM_I64"]:
        s = (int(a) + int(b)) & ((1 << 64) - 1)
        return s - (1 << 64) if s >= 1 << 63 else s
    if kind == C["TPLX_ACC_SUM_F64"]:
        return float(a) + float(b)
    if kind in (C["TPLX_ACC_MIN_I64"], C["TPLX_ACC_MIN_F64"]):
        return min(a, b)
    return max(a, b)


def _device_for(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def combine_aggregate(values: Sequence, kinds: Sequence[int], group=None) -> list:
    """All ranks contribute their partial aggregate (one value per accumulator); every rank gets the result
    combined in RANK ORDER — a fixed association, so f64 sums are reproducible run to run."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    is_f = [k in (C["TPLX_ACC_SUM_F64"], C["TPLX_ACC_MIN_F64"], C["TPLX_ACC_MAX_F64"]) for k in kinds]
    # ship the raw 64-bit patterns so that nothing is rounded in transit
    bits = np.array([np.float64(v).view(np.int64) if f else np.int64(v) for v, f in zip(values, is_f)], dtype=np.int64)
    t = torch.from_numpy(bits).to(_device_for(dist))
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    out = None
    for g in gathered:
        vals = g.cpu().numpy()
        cur = [float(np.int64(x).view(np.float64)) if f else int(x) for x, f in zip(vals, is_f)]
        out = cur if out is None else [_combine(k, a, b) for k, a, b in zip(kinds, out, cur)]
    return out


def allgather_arrays(arrays: Sequence[np.ndarray], group=None) -> List[List[np.ndarray]]:
    """Variable-length all_gather of several 1-D arrays: result[r][i] = rank r's arrays[i]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = _device_for(dist)
    lens = torch.tensor([len(a) for a in arrays], dtype=torch.int64, device=dev)
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    all_lens = [l.cpu().numpy() for l in all_lens]
    out: List[List[np.ndarray]] = [[None] * len(arrays) for _ in range(world)]
    for i, a in enumerate(arrays):
        mx = int(max(l[i] for l in all_lens))
        raw = np.zeros(max(mx, 1) * a.dtype.itemsize, dtype=np.uint8)
        raw[: a.nbytes] = np.ascontiguousarray(a).view(np.uint8)
        t = torch.from_numpy(raw).to(dev)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t, group=group)
        for r in range(world):
            n = int(all_lens[r][i])
            out[r][i] = gathered[r].cpu().numpy()[: n * a.dtype.itemsize].view(a.dtype).copy()
    return out


def exchange_hash_tables(stage, device: int, group=None):
    """aggregateByKey across ranks, on the device (tplx_gpu_stage_hash_exchange): afterwards every rank's table holds exactly
    the groups it owns (owner = hash(key) mod world); the union over ranks is the global result."""
    init_comm(device, group)
    stage.hash_exchange(device)
