"""Multi-GPU: one process per GPU, scaffolds sharded across ranks, ONE collective.

The path shards naturally: every rank counts the forward-strand k-mers of
its own scaffolds, the 87,380-counter tables (+ the genome-space scalar) are summed with a single
all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests), then every rank finalises the
tables redundantly and scores its own windows.  No other exchange exists on the data path.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_scaffolds(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of scaffolds to ranks, balanced by bases; scaffold order
    inside a rank is preserved (row order within a scaffold matters downstream: the HMM sees each
    scaffold's rows as one sequence, F:769)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(x) for x in out]


def global_genome_space(local_space: int, device=None, group=None) -> int:
    """Sum of totalLen - nnTotal over ranks (a property of the input, known at ingest time)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(local_space)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def make_allreduce(group=None):
    """Returns the hook for engine.Pipeline(allreduce=...): sums the forward counts over ranks in
    place -- the path's one collective.  The genome space passed in must already be the global one
    (global_genome_space) so the step needs no host round trip."""
    import torch.distributed as dist

    def hook(d_fwd, genome_space: int) -> int:
        dist.all_reduce(d_fwd, op=dist.ReduceOp.SUM, group=group)
        return genome_space

    return hook


def reduce_counts_cpu(tables: np.ndarray, genome_space: int, group=None) -> Tuple[np.ndarray, int]:
    """Same reduction on host arrays (gloo); used by the CPU world_size-2 tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    buf = torch.from_numpy(np.concatenate([tables.astype(np.int64), np.array([genome_space], np.int64)]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = buf.numpy()
    return out[:-1].astype(np.uint64), int(out[-1])
