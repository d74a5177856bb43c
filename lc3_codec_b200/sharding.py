"""Stream sharding for multi-GPU runs: streams are independent (src/decoder/lc3_decoder.rs:62-69), so rank r of N
owns the contiguous block of stream ids s with floor(s * N / total) == r.  No collective touches the data path."""
from __future__ import annotations


def shard_range(total_streams: int, rank: int, world: int) -> tuple[int, int]:
    """(first stream id, count) owned by `rank`; blocks differ by at most one stream and tile [0, total)."""
    if not (0 <= rank < world) or total_streams < 0:
        raise ValueError("bad rank/world/total")
    start = -(-rank * total_streams // world)            # ceil(rank * total / world)
    end = -(-(rank + 1) * total_streams // world)
    return start, end - start


def owner_of(stream: int, total_streams: int, world: int) -> int:
    return stream * world // total_streams


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device time of a multi-rank step is the slowest rank's (bench.py timing rule)."""
    if dist is None or not dist.is_initialized():
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
