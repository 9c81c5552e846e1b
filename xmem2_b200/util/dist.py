"""Multi-GPU plumbing for the stream-parallel mode: one process per GPU, independent video streams
per rank, NO collective on the data path.  The only communication is the timing reduction of the benchmark."""
from __future__ import annotations

import torch
import torch.distributed as dist


def streams_for_rank(n_streams: int, rank: int, world: int):
    """Indices of the video streams rank `rank` owns (round-robin; every stream owned by exactly one rank)."""
    return list(range(rank, n_streams, world))


def stream_seed(base_seed: int, stream_index: int) -> int:
    """BASELINE.json config 5: stream i uses seed base + 1000*i."""
    return base_seed + 1000 * stream_index


def max_over_ranks(value: float, device) -> float:
    """MAX all-reduce of a scalar (timing is always the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_throughput(frames_local: int, ms_local: float, device) -> float:
    """frames/s of the whole job = all ranks' frames / slowest rank's time."""
    total = sum_over_ranks(float(frames_local), device)
    ms = max_over_ranks(float(ms_local), device)
    return total / (ms * 1e-3)
