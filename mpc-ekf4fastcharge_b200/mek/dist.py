"""Multi-GPU plumbing: cells are independent (runMPC.m:83-112 has no coupling), so a population is
sharded by contiguous cell ranges, one process per GPU, with no data-path collective.  The only
exchange is the final reduction of the summary statistics (sum / min / max of ~16 scalars)."""
from __future__ import annotations

import numpy as np

from ._lib import STAT_NAMES

_MIN_KEYS = ("min_phise",)
_MAX_KEYS = ("max_V",)


def shard_range(B, rank, world):
    """Contiguous cell range [lo, hi) of `rank`; sizes differ by at most one cell.  (A noisy-sensor population passes `lo` as
    CellBatch(first_cell=lo): the noise sequence is keyed by the global cell index.)"""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_summaries(local, device=None, group=None):
    """All-reduce a mek_summary dict over the process group (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    dev = device if device is not None else "cpu"
    sums = torch.tensor([0.0 if (k in _MIN_KEYS or k in _MAX_KEYS) else float(local[k]) for k in STAT_NAMES], dtype=torch.float64, device=dev)
    mins = torch.tensor([float(local[k]) for k in _MIN_KEYS], dtype=torch.float64, device=dev)
    maxs = torch.tensor([float(local[k]) for k in _MAX_KEYS], dtype=torch.float64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    out = {k: float(v) for k, v in zip(STAT_NAMES, sums.cpu().tolist())}
    for k, v in zip(_MIN_KEYS, mins.cpu().tolist()):
        out[k] = float(v)
    for k, v in zip(_MAX_KEYS, maxs.cpu().tolist()):
        out[k] = float(v)
    return out


def max_over_ranks(value, device=None, group=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def merge_summaries_host(parts):
    """Reference reduction used by tests: combine per-shard summary dicts on the host."""
    out = {k: 0.0 for k in STAT_NAMES}
    out["min_phise"] = np.inf
    out["max_V"] = -np.inf
    for p in parts:
        for k in STAT_NAMES:
            if k in _MIN_KEYS:
                out[k] = min(out[k], p[k])
            elif k in _MAX_KEYS:
                out[k] = max(out[k], p[k])
            else:
                out[k] += p[k]
    return out
