"""Multi-GPU plumbing of the hot path: videos shard one per rank, one collective at the very end.

The reference has no distributed code; tracker state is reset per video
(/root/reference/tracklab/engine/offline.py:11-13), so videos are independent units: video ``i`` goes to rank
``i % world`` and runs detect -> associate locally with zero communication. The only collective is one
``all_gather`` of a fixed-size per-video metric vector (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

METRIC_FIELDS = ("frames", "detections", "rows", "track_ids", "ms_per_step")


def shard_videos(n_videos: int, rank: int, world: int) -> list[int]:
    """Indices of the videos rank ``rank`` owns (round-robin, like SURVEY.md §8e)."""
    return list(range(rank, n_videos, world))


def gather_video_metrics(local: torch.Tensor) -> torch.Tensor:
    """local: float64 [n_local_videos, len(METRIC_FIELDS)] (same n on every rank) -> [world, n_local, F] on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local[None].clone()
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local.contiguous())
    return torch.stack(out)


def max_over_ranks(ms: float, device) -> float:
    """Step time of the job = slowest rank (device-timed on each rank)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
