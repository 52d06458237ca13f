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


class OverlappedAllReduce:
    """The step's gradient all-reduce in two buckets (``taco_dp_bucket``): the decoder / post-net / linear gradients (the tail of
    the flat buffer, ~62 % of it) are reduced on a communication stream as soon as the decoder's backward pass has produced
    them - beside the encoder's backward pass -, the embedding / speaker / encoder gradients once the backward pass is complete.
    Call it where ``allreduce_sum_`` would be called: right after ``Engine.backward()`` has been enqueued.  Returns 1/world."""

    def __init__(self, engine):
        import ctypes as C
        from . import capi
        self.engine, self._C, self._capi = engine, C, capi
        off, num = C.c_int64(), C.c_int64()
        capi.check(engine.lib.taco_dp_bucket(engine._h, 0, C.byref(off), C.byref(num)))
        self.early = (int(off.value), int(num.value))
        self.comm = torch.cuda.Stream(device=engine.dev) if world() > 1 else None

    def __call__(self, flat: torch.Tensor) -> float:
        w = world()
        if w == 1:
            return 1.0
        eng = self.engine
        off, num = self.early
        main = torch.cuda.current_stream(eng.dev)
        if num > 0:
            self._capi.check(eng.lib.taco_dp_wait_bucket(eng._h, 0, self.comm.cuda_stream))
            with torch.cuda.stream(self.comm):
                dist.all_reduce(flat[off:off + num], op=dist.ReduceOp.SUM)
        if off > 0:
            dist.all_reduce(flat[:off], op=dist.ReduceOp.SUM)      # current stream: behind the whole backward pass
        if num > 0:
            main.wait_stream(self.comm)
        return 1.0 / w
