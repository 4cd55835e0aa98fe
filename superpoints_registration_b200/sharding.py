"""Multi-GPU sharding of the registration path: pairs are independent, so ranks take contiguous slices of the
pair list and the only collective is the final gather of the (B_local, 3, 4) poses.

The reference's own multi-GPU story is DistributedSampler + DDP for training (train.py:57-64,
data_loaders/__init__.py:76); for inference that reduces to exactly this: disjoint pair subsets per rank.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(costs: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous slices [lo, hi) per rank, balanced by cumulative cost (points per pair); every pair is
    assigned exactly once, ranks may be empty when there are fewer pairs than ranks."""
    n = len(costs)
    total = float(sum(costs))
    bounds, lo, acc = [], 0, 0.0
    for r in range(world_size):
        target = total * (r + 1) / world_size
        hi = lo
        while hi < n and (acc + costs[hi] / 2.0 <= target or hi == lo) and (n - hi) > (world_size - 1 - r):
            acc += costs[hi]
            hi += 1
        if r == world_size - 1:
            hi = n
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_batch(batch: dict, rank: int, world_size: int) -> Tuple[dict, List[Tuple[int, int]]]:
    """Slice a collate_pair-style batch ({'src_xyz': [...], 'tgt_xyz': [...], ...}) for this rank."""
    costs = [int(s.shape[0] + t.shape[0]) for s, t in zip(batch["src_xyz"], batch["tgt_xyz"])]
    bounds = shard_bounds(costs, world_size)
    lo, hi = bounds[rank]
    local = {k: (v[lo:hi] if isinstance(v, (list, tuple)) or torch.is_tensor(v) else v) for k, v in batch.items()}
    return local, bounds


def gather_poses(local_pose: torch.Tensor, bounds: Sequence[Tuple[int, int]], group=None) -> torch.Tensor:
    """all_gather of per-rank poses (possibly different counts) into the global (B, 3, 4) tensor, pair order kept."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_pose
    world = dist.get_world_size(group)
    counts = [hi - lo for lo, hi in bounds]
    width = max(max(counts), 1)
    padded = torch.zeros((width, 3, 4), dtype=local_pose.dtype, device=local_pose.device)
    padded[:local_pose.shape[0]] = local_pose
    out = torch.empty((world * width, 3, 4), dtype=local_pose.dtype, device=local_pose.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view(world, width, 3, 4)
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)
