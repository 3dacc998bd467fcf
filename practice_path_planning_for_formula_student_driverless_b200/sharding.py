"""Multi-GPU plumbing: one process per GPU, problems sharded in contiguous blocks, no hot-path collective.

Every (track, Config) problem is independent (the reference solvers are pure functions of their
arguments and Config, main.cpp:683-764 / 905-1052), so rank r of G owns problems
[floor(r*B/G), floor((r+1)*B/G)) and the only exchange is the final gather of per-problem lap times
(8 bytes per problem) -- or a tiny all-reduce(min) when the consumer is "best Config of a sweep".
torch.distributed is used for exactly that (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_bounds(n_problems: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of the problem index owned by `rank` (problem p -> rank floor(p*G/B))."""
    if world_size <= 0 or not (0 <= rank < world_size) or n_problems < 0:
        raise ValueError("bad shard request")
    lo = (n_problems * rank) // world_size
    hi = (n_problems * (rank + 1)) // world_size
    return lo, hi


def owner_of(problem: int, n_problems: int, world_size: int) -> int:
    """Inverse of shard_bounds."""
    r = min(world_size - 1, (problem * world_size) // max(1, n_problems))
    while problem < shard_bounds(n_problems, world_size, r)[0]:
        r -= 1
    while problem >= shard_bounds(n_problems, world_size, r)[1]:
        r += 1
    return r


def gather_lap_times(local_laps, n_problems: int, device=None):
    """Final gather: every rank ends with the lap time of every problem, in global problem order."""
    import torch
    import torch.distributed as dist

    # a torch tensor (e.g. DeviceBatch.device_tensor("lap_time"), already on the GPU) is gathered where it lies
    local = local_laps.to(torch.float64).contiguous() if isinstance(local_laps, torch.Tensor) else \
        torch.as_tensor(np.asarray(local_laps, dtype=np.float64))
    if device is not None:
        local = local.to(device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.cpu().numpy()
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(n_problems, world, rank)
    assert local.numel() == hi - lo, "local lap count does not match this rank's shard"
    width = max(shard_bounds(n_problems, world, r)[1] - shard_bounds(n_problems, world, r)[0] for r in range(world))
    pad = torch.full((width,), float("inf"), dtype=torch.float64, device=local.device)
    pad[:local.numel()] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_problems, world, r)
        parts.append(out[r][: b - a].cpu().numpy())
    return np.concatenate(parts) if parts else np.zeros(0)


def gather_rasters(local_rows, n_problems: int, rows_per_problem: int):
    """Final gather of per-sample rasters (e.g. the raceline xy rows) of equally long problems: every rank ends with a
    tensor of all problems' rows in global problem order.  `local_rows`: torch tensor (local problems * rows_per_problem,
    ...) on this rank's device; one all_gather over NCCL (NVLink / NVSwitch) or gloo."""
    import torch
    import torch.distributed as dist

    local = local_rows.contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(n_problems, world, rank)
    assert local.shape[0] == (hi - lo) * rows_per_problem, "local rows do not match this rank's shard"
    width = max(shard_bounds(n_problems, world, r)[1] - shard_bounds(n_problems, world, r)[0] for r in range(world))
    tail = tuple(local.shape[1:])
    pad = torch.zeros((width * rows_per_problem,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world, width * rows_per_problem) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view((world * width * rows_per_problem,) + tail), pad)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_problems, world, r)
        parts.append(out[r, : (b - a) * rows_per_problem])
    return torch.cat(parts, dim=0)


def best_of_sweep(local_laps, n_problems: int, device=None) -> Tuple[int, float]:
    """arg-min lap over a sharded Config sweep: one all-reduce(min) on (lap, global index) pairs."""
    import torch
    import torch.distributed as dist

    if isinstance(local_laps, torch.Tensor):
        device = local_laps.device if device is None else device
        local_laps = local_laps.detach().cpu().numpy()
    laps = np.asarray(local_laps, dtype=np.float64)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world, rank = (dist.get_world_size(), dist.get_rank()) if distributed else (1, 0)
    lo, _ = shard_bounds(n_problems, world, rank)
    if laps.size:
        i = int(np.argmin(laps))
        best = (float(laps[i]), lo + i)
    else:
        best = (float("inf"), n_problems)
    if not distributed:
        return best[1], best[0]
    t = torch.tensor([best[0]], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    # ties: smallest global index among the ranks that hold the minimum
    cand = torch.tensor([best[1] if best[0] == t.item() else n_problems], dtype=torch.int64, device=device)
    dist.all_reduce(cand, op=dist.ReduceOp.MIN)
    return int(cand.item()), float(t.item())
