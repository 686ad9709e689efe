"""One-process-per-GPU plumbing for the sample-sharded hot path.

Rows A, B, C, F of the hot path shard by sample with NO data-path collective: rank r owns samples
[r * B_local, (r + 1) * B_local) and nothing is exchanged.  The only cross-rank operations are the ones the
measurement contract needs (max of device times, sums of processed units).  Works on NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous, balanced shard of n units: the first n % world ranks get one extra unit."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Distinct, reproducible synthetic-input stream per rank (weak scaling: every rank makes its own batch)."""
    return seed + 1000003 * rank


def _reduce(value: float, op, device=None) -> float:
    rank, ws = world()
    if ws == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value: float, device=None) -> float:
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value: float, device=None) -> float:
    return _reduce(value, dist.ReduceOp.SUM, device)


def job_throughput(units_this_rank: float, elapsed_ms_this_rank: float, device=None) -> float:
    """Whole-job units/s: all units processed by all ranks divided by the slowest rank's time."""
    total = sum_over_ranks(units_this_rank, device)
    slowest = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (slowest * 1e-3)


def bucket_ranges(lo: int, hi: int, bucket_elems: int):
    """[lo, hi) cut into buckets of at most bucket_elems elements, LAST bucket first: the order in which the backward
    pass finishes the flat gradient buffer (model tail first)."""
    out, pos = [], hi
    step = max(1, int(bucket_elems))
    while pos > lo:
        a = max(lo, pos - step)
        out.append((a, pos))
        pos = a
    return out


def allreduce_range(flat_grad: torch.Tensor, lo: int, hi: int, bucket_elems: int, group=None):
    """Sum-all-reduce flat_grad[lo:hi) in buckets (the data-parallel gradient exchange)."""
    for a, b in bucket_ranges(lo, hi, bucket_elems):
        dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM, group=group)


class SectionCoalescer:
    """Groups the gradient sections a backward pass announces (each [lo, hi) of the flat buffer, arriving from the tail of
    the buffer towards its head) into exchanges of at least `min_elems` elements: `add` returns the ranges to all-reduce
    now -- nothing while less than `min_elems` are pending, everything pending once the head (lo == 0) is reached.
    A section that is not adjacent to the pending range flushes it first, so no element is ever exchanged twice or left out."""

    def __init__(self, min_elems: int):
        self.min_elems = int(min_elems)
        self.lo = self.hi = None

    def add(self, lo: int, hi: int):
        out = []
        if hi <= lo:
            return out
        if self.hi is not None and self.lo != hi:
            out.append((self.lo, self.hi))
            self.lo = self.hi = None
        if self.hi is None:
            self.hi = hi
        self.lo = lo
        if lo == 0 or self.hi - lo >= self.min_elems:
            out.append((self.lo, self.hi))
            self.lo = self.hi = None
        return out

    def flush(self):
        out = [] if self.hi is None else [(self.lo, self.hi)]
        self.lo = self.hi = None
        return out
