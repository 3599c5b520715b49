"""View-parallel multi-GPU plumbing (one process per GPU, ``torch.distributed``).

The reference is single-GPU and renders one view per step (GaussianPointTrainer.py:120-166).  The path
shards naturally by VIEW: every rank holds a full replica of the scene, view ``i`` is
rendered by rank ``i mod R``, inference needs no communication, and training needs exactly one exchange
step per optimiser step -- the sum of the dense per-Gaussian gradients (N,3)+(N,56) over ranks, one
NCCL all-reduce over NVLink (gloo on CPU in the tests).
"""
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world_size: int) -> List[int]:
    """Indices of the views rank ``rank`` renders: i with i mod world_size == rank."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, num_views, world_size))


def _aliases(buffer: torch.Tensor, grads) -> bool:
    """True if every gradient tensor is a view into ``buffer``'s allocation (then one collective covers all)."""
    lo = buffer.data_ptr()
    hi = lo + buffer.numel() * buffer.element_size()
    return all(g is None or (g.is_contiguous() and lo <= g.data_ptr() and
                             g.data_ptr() + g.numel() * g.element_size() <= hi) for g in grads)


def exchange_gradients(grads: Iterable[Optional[torch.Tensor]], group=None, average: bool = False,
                       async_op: bool = False, fused_buffer: Optional[torch.Tensor] = None):
    """Sum (or average) the dense gradient tensors over all ranks, in place.

    ``fused_buffer``: the operator's ``last_gradient_buffer``; when the gradients are views into it (autograd
    hands the operator's outputs to ``.grad`` without copying) the exchange is ONE all-reduce instead of one
    per tensor.

    Returns the list of work handles when ``async_op`` (so the exchange can overlap the next view's
    forward on another stream), else ``None``.  A no-op when torch.distributed is not initialised or
    the world has a single rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [] if async_op else None
    handles = []
    world = dist.get_world_size(group)
    grads = list(grads)
    if fused_buffer is not None and _aliases(fused_buffer, grads):
        grads = [fused_buffer]
    for g in grads:
        if g is None:
            continue
        h = dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            handles.append(h)
        if average:
            if async_op:
                h.wait()
            g.div_(world)
    return handles if async_op else None


def render_views(op, make_input, view_ids: Sequence[int]):
    """Inference helper: render this rank's shard of views (no communication). ``make_input(i)`` builds the
    ``GaussianPointCloudRasterisationInput`` of view i."""
    out = {}
    with torch.no_grad():
        for i in view_ids:
            image, depth, count = op(make_input(i))
            out[i] = (image, depth, count)
    return out
