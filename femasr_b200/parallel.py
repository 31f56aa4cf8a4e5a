"""Data-parallel sharding of the hot path across the GPUs of one node (one process per GPU).

Images (and test_tile tiles) are independent - GroupNorm is per-sample, LayerNorm per-token, attention
per-window, the VQ per-pixel against a replicated codebook - so the batch is split
contiguously, every rank runs the whole path on its shard with replicated weights, and the only
collective is ONE all-gather of the output shards (NCCL over NVLink on the GPU box; gloo in the CPU tests).
The reference has no multi-GPU inference path (basicsr/models/femasr_model.py:229-232).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

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


def tile_classes(height: int, width: int, tile_size: int, tile_pad: int) -> Dict[Tuple[int, int], List[dict]]:
    """The tiles of FeMaSRNet.test_tile (femasr_arch.py:387-447) grouped by the shape of their padded input window:
    tiles of one class can be stacked into one batch.  Deterministic order (row-major inside a class)."""
    from .net import tile_plan                      # pure Python (no CUDA needed to plan)
    classes: Dict[Tuple[int, int], List[dict]] = {}
    for t in tile_plan(height, width, tile_size, tile_pad):
        py0, py1, px0, px1 = t["in"]
        classes.setdefault((py1 - py0, px1 - px0), []).append(t)
    return classes


def sharded_test_tile(run: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, scale: int, tile_size: int,
                      tile_pad: int, rank: int, world: int, group: Optional[dist.ProcessGroup] = None,
                      max_batch: int = 64) -> torch.Tensor:
    """test_tile with the TILE LIST sharded over the ranks: inside every same-shape class the tiles are
    dealt round-robin, each rank runs `run` (e.g. FeMaSRNet.test) on stacks of its tiles, one all-gather per class
    brings the SR tiles to every rank, and every rank pastes the full [B,3,sH,sW] result.  `run` maps
    [n,3,th,tw] -> [n,3,s*th,s*tw]; tiles are independent on this path, so stacking does not change their values."""
    B, C, H, W = x.shape
    s = scale
    out = x.new_zeros((B, C, H * s, W * s))
    for (th, tw), tiles in tile_classes(H, W, tile_size, tile_pad).items():
        mine = tiles[rank::world]
        counts = [len(tiles[r::world]) * B for r in range(world)]
        parts = []
        per = max(1, max_batch // B)
        for i in range(0, len(mine), per):
            chunk = mine[i:i + per]
            stack = torch.cat([x[:, :, t["in"][0]:t["in"][1], t["in"][2]:t["in"][3]] for t in chunk], 0)
            parts.append(run(stack.contiguous()))
        local = torch.cat(parts, 0) if parts else x.new_zeros((0, C, th * s, tw * s))
        full = all_gather_outputs(local, counts, group)          # rank-major: rank 0's tiles, rank 1's tiles, ...
        order = [t for r in range(world) for t in tiles[r::world]]
        for k, t in enumerate(order):
            y0, y1, x0, x1 = t["out"]
            cy, cx = t["crop"]
            sr = full[k * B:(k + 1) * B]
            out[:, :, y0 * s:y1 * s, x0 * s:x1 * s] = sr[:, :, cy * s:(cy + y1 - y0) * s, cx * s:(cx + x1 - x0) * s]
    return out
