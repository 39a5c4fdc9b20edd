"""Data-parallel plumbing for the hot path: one process per GPU, replicated weights,
the image batch sharded by rank, and exactly one exchange per training step -- a mean all-reduce of the
gradients (the reference trains single-process, src/train/train.py:164-178; this is the `torch.distributed`
equivalent of running it under DDP).

The gradients of the encoder/decoder already live in ONE flat fp32 buffer (`runtime.FlatParams.g32`), so
they go to NCCL as in-place all-reduces with no flatten/unflatten copies (with the hand-scheduled runtime the
class / box heads live in that buffer too); tensors outside it, if any, are coalesced into one more call.  Nothing here touches the device
directly, which is why the same code runs under `gloo` on CPU tensors in tests/test_dataparallel_gloo.py.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> range:
    """Images [r*b, (r+1)*b) of a global batch go to rank r; the batch must divide evenly."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not divide over {world} ranks")
    b = global_batch // world
    return range(rank * b, (rank + 1) * b)


def allreduce_mean_(flat_buffers: Sequence[torch.Tensor], loose: Iterable[Optional[torch.Tensor]] = (),
                    world: Optional[int] = None, group=None) -> None:
    """In-place mean over ranks of every tensor in `flat_buffers` (each reduced as is, one collective per
    buffer) and of the `loose` tensors (coalesced into a single extra collective).  No-op for world == 1."""
    world = dist.get_world_size(group) if world is None else world
    if world == 1:
        return
    inv = 1.0 / world
    for buf in flat_buffers:
        dist.all_reduce(buf, group=group)
        buf.mul_(inv)
    rest: List[torch.Tensor] = [t for t in loose if t is not None]
    if rest:
        flat = torch._utils._flatten_dense_tensors(rest)
        dist.all_reduce(flat, group=group)
        flat.mul_(inv)
        torch._foreach_copy_(rest, list(torch._utils._unflatten_dense_tensors(flat, rest)))


def grads_of(model: torch.nn.Module):
    """(flat buffers, loose grads) of a TransformerHalf-like module: the runtime's flat gradient buffer when the
    hand-scheduled runtime is active, plus the .grad of every parameter that does not live inside it."""
    rt = getattr(model, "_rt", None)
    if rt is None:
        return [], [p.grad for p in model.parameters()]
    flat = rt.P.g32
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
    loose = [p.grad for p in model.parameters() if p.grad is not None and not (lo <= p.grad.data_ptr() < hi)]
    if getattr(rt.P, "world", 1) > 1:  # the runtime already exchanged its flat buffer inside backward()
        return [], loose
    return [flat], loose
