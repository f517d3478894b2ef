"""Multi-GPU plumbing: sequences are independent, so the path shards with no
data-path collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to
broadcast the parsed config and to gather per-sequence trajectories/states on rank 0."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_sequences(n_total: int, rank: int, world: int) -> List[int]:
    """Contiguous block partition of sequence ids; the first ``n_total % world`` ranks get one more."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return list(range(lo, lo + base + (1 if rank < rem else 0)))


def gather_states(local: torch.Tensor, n_total: int, rank: int, world: int) -> Optional[torch.Tensor]:
    """local: [n_local, 17] (t, q, p, v, bg, ba) -> rank 0 gets [n_total, 17] in sequence order, others None."""
    counts = [len(shard_sequences(n_total, r, world)) for r in range(world)]
    mx = max(counts)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != 0:
        return None
    return torch.cat([bufs[r][:counts[r]] for r in range(world)], 0)


def broadcast_config_bytes(blob: Optional[bytes], rank: int) -> bytes:
    """Rank 0's config file contents to every rank (one tiny broadcast at start-up)."""
    obj = [blob]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


def share_sequences(local: list, n_total: int, rank: int, world: int, device=None) -> list:
    """Every rank generated the sequences ``shard_sequences(n_total, rank, world)`` of one pool; after this call every
    rank holds the whole pool, in pool order.  One all_gather per field (images dominate: n_total x F x H x W bytes,
    over NVLink when ``device`` is a GPU); fields are padded to the longest shard / longest IMU record.
    Used by bench.py so that host-side image synthesis does not grow with the number of GPUs."""
    import numpy as np
    from types import SimpleNamespace
    counts = [len(shard_sequences(n_total, r, world)) for r in range(world)]
    mx = max(counts)
    dev = torch.device("cpu") if device is None else device
    n_imu = torch.tensor([max([len(q.imu) for q in local] + [0])], dtype=torch.int64, device=dev)
    dist.all_reduce(n_imu, op=dist.ReduceOp.MAX)
    n_imu = int(n_imu.item())
    fields = ["images", "img_t", "imu", "gt_p", "gt_q", "gt_v", "gyro_bias", "acc_bias"]
    out = {}
    lens = torch.zeros(mx, dtype=torch.int64, device=dev)
    for i, q in enumerate(local):
        lens[i] = len(q.imu)
    lens_all = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(lens_all, lens)
    for f in fields:
        arrs = []
        for q in local:
            a = np.asarray(getattr(q, f))
            if f == "imu":
                pad = np.zeros((n_imu, a.shape[1]), a.dtype); pad[:len(a)] = a; a = pad
            arrs.append(a)
        proto = arrs[0]
        buf = np.zeros((mx,) + proto.shape, proto.dtype)
        for i, a in enumerate(arrs):
            buf[i] = a
        t = torch.from_numpy(buf).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        out[f] = [parts[r][:counts[r]].cpu().numpy() for r in range(world)]
    pool = []
    for r in range(world):
        for i in range(counts[r]):
            d = {f: out[f][r][i] for f in fields}
            d["imu"] = d["imu"][:int(lens_all[r][i].item())]
            pool.append(SimpleNamespace(**d))
    return pool
