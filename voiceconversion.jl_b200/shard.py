"""Sharding of a batch over the GPUs of one box (one process per GPU).

Every unit of the path is independent -- frames for ``vc(::GMMMap)``, utterances / chunks for
``vc(::TrajectoryGMMMap)``, pairs for DTW -- so ranks take contiguous slices of the batch, replicate
the (small) model, and never exchange data while computing.  The only collective is an optional
gather of the results (NCCL when the tensors live on the GPU, gloo on the CPU in tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def partition_contiguous(costs: Sequence[float], nparts: int) -> List[Tuple[int, int]]:
    """Splits units 0..n-1 into ``nparts`` contiguous ranges with near-equal total cost (frames
    per utterance, S*T per DTW pair).  Returns [(begin, end)] per part; parts may be empty."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if nparts < 1:
        raise ValueError("nparts must be >= 1")
    csum = np.concatenate([[0.0], np.cumsum(costs)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, nparts):
        target = total * r / nparts
        # first index whose prefix cost reaches the target, never moving backwards
        idx = int(np.searchsorted(csum, target, side="left"))
        idx = min(max(idx, bounds[-1]), n)
        # pick the closer of idx-1 / idx
        if idx > bounds[-1] and abs(csum[idx - 1] - target) <= abs(csum[idx] - target):
            idx -= 1
        bounds.append(idx)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(nparts)]


def frame_range(T: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous frame range [begin, end) of ``rank`` for frame-by-frame conversion."""
    return (T * rank) // world, (T * (rank + 1)) // world


def shard_ragged(offsets: Sequence[int], rank: int, world: int, cost_fn: Callable[[np.ndarray], np.ndarray] = None):
    """For a ragged batch described by ``offsets`` (n+1): the unit range of ``rank`` and the
    rebased offsets of its slice.  Units are balanced by length (or ``cost_fn(lengths)``)."""
    off = np.asarray(offsets, dtype=np.int64)
    lens = np.diff(off)
    costs = cost_fn(lens) if cost_fn is not None else lens
    b, e = partition_contiguous(costs, world)[rank]
    return (b, e), off[b:e + 1] - off[b]


def gather_frames(local, dist=None, dst: int = 0, sizes: Sequence[int] = None):
    """Optional result gather to rank ``dst`` along the frame axis (axis 0 of a frame-major torch
    tensor).  Every other rank sends its shard straight into its slice of ``dst``'s preallocated
    batch (grouped ``isend``/``irecv``: ncclSend/ncclRecv over NVLink for CUDA tensors, gloo for CPU
    tensors), so nothing is padded or copied twice; shards may have different lengths.
    ``sizes`` (frames per rank) skips the size exchange when the caller knows the partition
    (``shard_ragged`` is deterministic).  Returns the concatenated tensor on ``dst``, None elsewhere."""
    import torch
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    world, rank = dist.get_world_size(), dist.get_rank()
    if sizes is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        got = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(got, n)
        sizes = [int(s.item()) for s in got]
    sizes = [int(x) for x in sizes]
    if len(sizes) != world or sizes[rank] != local.shape[0]:
        raise ValueError("sizes must list the frame count of every rank")
    if rank != dst:
        if sizes[rank] > 0:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst)]):
                w.wait()
        return None
    starts = np.concatenate([[0], np.cumsum(sizes)])
    out = torch.empty((int(starts[-1]),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    ops = [dist.P2POp(dist.irecv, out[int(starts[r]):int(starts[r + 1])], r)
           for r in range(world) if r != dst and sizes[r] > 0]
    works = dist.batch_isend_irecv(ops) if ops else []
    out[int(starts[dst]):int(starts[dst + 1])].copy_(local)
    for w in works:
        w.wait()
    return out
