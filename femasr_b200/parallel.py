"""Data-parallel sharding of the hot path across the GPUs of one node (one process per GPU).

Images (and test_tile tiles) are independent - GroupNorm is per-sample, LayerNorm per-token, attention
per-window, the VQ per-pixel against a replicated codebook - so the batch is split
contiguously, every rank runs the whole path on its shard with replicated weights, and the only
collective is ONE all-gather of the output shards (NCCL over NVLink on the GPU box; gloo in the CPU tests).
The reference has no multi-GPU inference path (basicsr/models/femasr_model.py:229-232).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank's images; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_counts(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def all_gather_outputs(local: torch.Tensor, counts: List[int], group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Concatenate the per-rank output shards (rank order) on every rank with a single all-gather.
    Equal shards use all_gather_into_tensor directly into the result; ragged shards are padded to the
    largest shard for the collective and trimmed afterwards."""
    world = len(counts)
    if world == 1:
        return local
    tail = tuple(local.shape[1:])
    mx = max(counts)
    if mx == 0:
        return local.new_zeros((0,) + tail)
    if all(c == mx for c in counts):
        out = local.new_empty((world * mx,) + tail)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = local.new_zeros((mx,) + tail)
    padded[: local.shape[0]] = local
    buf = local.new_empty((world * mx,) + tail)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + c] for r, c in enumerate(counts)], 0)


def sharded_forward(run: Callable[[torch.Tensor], torch.Tensor], x_global: torch.Tensor, rank: int, world: int,
                    group: Optional[dist.ProcessGroup] = None, out_shape_fn: Optional[Callable] = None) -> torch.Tensor:
    """Run `run` (e.g. FeMaSRNet.test) on this rank's contiguous slice of x_global [N,3,H,W] and return
    the gathered [N,3,sH,sW] result on every rank."""
    n = x_global.shape[0]
    a, b = shard_range(n, rank, world)
    counts = shard_counts(n, world)
    if b > a:
        local = run(x_global[a:b])
    else:
        if out_shape_fn is None:
            raise ValueError("an empty shard needs out_shape_fn to size its (empty) output")
        local = x_global.new_zeros((0,) + tuple(out_shape_fn(x_global.shape)[1:]))
    return all_gather_outputs(local, counts, group)
