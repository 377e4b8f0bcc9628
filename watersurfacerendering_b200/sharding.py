"""Multi-GPU partitioning of the hot path: one process per GPU, NO data-path collective.

ComputeWaves(t) is a pure function of (h0, t, lambda) (reference: WSTessendorf.cpp:284-441 reads only
m_BaseWaveHeights / m_WaveVectors), so animation frames and independent tiles shard round-robin:
item i -> rank i mod P.  Only bookkeeping (amplitudes, timings) is ever gathered, over the process group
torch.distributed provides (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def shard_indices(n_items: int, rank: int, world: int) -> np.ndarray:
    """Global indices of the tile-frames rank `rank` owns (round-robin)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_items, world, dtype=np.int64)


def shard_counts(n_items: int, world: int) -> List[int]:
    return [len(range(r, n_items, world)) for r in range(world)]


def gather_in_global_order(local_vals: Sequence[float], n_items: int, rank: int, world: int,
                           group=None) -> Optional[np.ndarray]:
    """All ranks contribute the per-item scalars they own (e.g. amplitudes A); every rank gets the
    full vector in global item order.  Bookkeeping only - the maps never leave their GPU."""
    import torch
    import torch.distributed as dist

    counts = shard_counts(n_items, world)
    width = max(counts) if counts else 0
    loc = torch.full((width,), float("nan"), dtype=torch.float32)
    lv = torch.as_tensor(np.asarray(local_vals, np.float32))
    loc[: lv.numel()] = lv
    if world == 1 or not dist.is_initialized():
        bufs = [loc]
    else:
        dev = None
        if dist.get_backend(group) == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            loc = loc.to(dev)
        bufs = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(bufs, loc, group=group)
        bufs = [b.cpu() for b in bufs]
    out = np.full(n_items, np.nan, np.float32)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][: len(idx)].numpy()
    return out
