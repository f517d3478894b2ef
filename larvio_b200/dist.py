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
