"""Data-parallel plumbing: one process per GPU, ONE gradient all-reduce per training step.

The reference is single-device; this is the only collective the path has.  Replicas keep their own batch-norm batch
statistics (the reference's semantics at batch=32 per device); gradients are summed over ranks and the 1/world scale is
folded into the clip+Adam kernel, so every rank applies the identical update and replicas stay bit-identical."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum_(flat: torch.Tensor) -> float:
    """In-place SUM all-reduce of the flat gradient buffer; returns the scale (1/world) the optimizer must apply."""
    w = world()
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return 1.0 / w


def average_bn_state_(state: torch.Tensor) -> None:
    """Reconcile batch-norm moving statistics across replicas (tiny: 9 376 floats); call at checkpoint time."""
    w = world()
    if w > 1:
        dist.all_reduce(state, op=dist.ReduceOp.SUM)
        state.mul_(1.0 / w)


def shard_rows(n_rows: int, rank: int, world_size: int):
    """Replica sharding of independent utterances for batched inference (no collective): [begin, end) of this rank."""
    per = (n_rows + world_size - 1) // world_size
    return min(n_rows, rank * per), min(n_rows, (rank + 1) * per)
