"""Multi-GPU plumbing: the image stream shards embarrassingly, one process per GPU.

Every image's score depends only on that image and on replicated weights / bank
(``utils/detection_util.py:225-248`` has no cross-sample op), so rank ``r`` of ``W`` scores the
contiguous slice ``[r * ceil(N/W), min(N, (r+1) * ceil(N/W)))`` of each stream and a single
all-gather of the padded per-rank score vectors collates them.  There is no
collective on the data path; the gather moves <= 25 KB per rank per stream.

``torch.distributed`` is the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_len", "gather_scores", "world"]


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_len(n: int, world_size: int) -> int:
    return -(-int(n) // int(world_size))


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open slice of an ``n``-image stream owned by ``rank`` (possibly empty for tail ranks)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    per = shard_len(n, world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_scores(local, n: int, group=None, device: Optional[torch.device] = None) -> np.ndarray:
    """All-gather the per-rank score slices of an ``n``-image stream; every rank gets float32 ``[n]``.

    ``local`` (numpy or tensor) holds this rank's ``shard_bounds`` slice.  Slices are padded to
    ``ceil(n / W)`` so the collective is a plain equal-count all-gather, and the result is trimmed
    to ``n`` -- the same trim the reference applies to its own padded tail (``:249``).
    """
    rank, W = world()
    t = torch.as_tensor(local, dtype=torch.float32).reshape(-1)
    lo, hi = shard_bounds(n, rank, W)
    if t.numel() != hi - lo:
        raise ValueError(f"rank {rank} holds {t.numel()} scores, expected {hi - lo} for n={n}, world={W}")
    if W == 1:
        return t.detach().cpu().numpy().astype(np.float32, copy=True)
    if device is None:
        backend = dist.get_backend(group)
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    per = shard_len(n, W)
    send = torch.zeros((per,), dtype=torch.float32, device=device)
    send[: t.numel()] = t.to(device)
    recv = torch.empty((W * per,), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n].cpu().numpy().astype(np.float32, copy=True)
