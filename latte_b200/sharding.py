"""Data-parallel plumbing for sampling replicas (the only parallelism the reference has — SURVEY.md §2.4).

Each video's denoising trajectory is independent, so ranks take disjoint videos exactly as
`sample/sample_ddp.py:116-125,171-173` does (global index = i * world + rank) and no collective sits on
the data path.  `gather_frames` is the optional end-of-job uint8 gather (a NEW feature — the reference
has every rank write its own files); `max_over_ranks` is the timing reduction bench.py uses."""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def rank_seed(global_seed: int, rank: int, world_size: int) -> int:
    """sample_ddp.py:63-65: seed = global_seed * world_size + rank."""
    return global_seed * world_size + rank


def plan_iterations(num_samples: int, per_proc_batch: int, world_size: int) -> tuple[int, int]:
    """sample_ddp.py:116-125: pad the job to a multiple of the global batch; -> (total_samples, iterations per rank)."""
    global_batch = per_proc_batch * world_size
    total = int(math.ceil(num_samples / global_batch) * global_batch)
    return total, total // global_batch


def video_indices(iteration: int, per_proc_batch: int, rank: int, world_size: int) -> list[int]:
    """sample_ddp.py:171-173: sample i of this rank's batch in iteration k has global index
    i * world + rank + k * global_batch."""
    total_before = iteration * per_proc_batch * world_size
    return [i * world_size + rank + total_before for i in range(per_proc_batch)]


def max_over_ranks(value: float, device=None) -> float:
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frames(frames_u8: torch.Tensor) -> torch.Tensor | None:
    """all_gather this rank's decoded uint8 frames [n, F, H, W, 3] to every rank, interleaved back into global
    video order (inverse of `video_indices`).  Returns [n * world, F, H, W, 3]."""
    rank, ws = world()
    if ws == 1:
        return frames_u8
    parts = [torch.empty_like(frames_u8) for _ in range(ws)]
    dist.all_gather(parts, frames_u8.contiguous())
    stacked = torch.stack(parts, dim=1)  # [n, world, ...]: global index = i * world + r
    return stacked.reshape(-1, *frames_u8.shape[1:])
