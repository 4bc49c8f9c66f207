"""View-sharded data parallelism for the render path.

One process per GPU, full replica of the Gaussian table; each rank renders its own
views.  The reference defines the semantics through ``batch_size=B`` on one GPU:
parameter gradients are the MEAN over the views of a batch (the loss is a mean over
the stacked batch, pointrix/model/loss.py:27-46), the densification statistic is the
SUM over views of each ``ndc.grad`` (pointrix/controller/gs.py:274-278), ``radii`` is
the MAX and ``visibility`` the ANY over views (pointrix/model/renderer/msplat.py:211-212).
``world_size`` ranks x 1 view therefore equals a reference batch of ``world_size``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int, step: int = 0) -> List[int]:
    """Views of one step: view_id % world == rank, with a shared rotation by ``step``."""
    return [v for v in range(num_views) if (v + step) % world == rank]


def _coalesce(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Tensors that tile one contiguous range of a shared storage (the fused backward's flat
    gradient buffer) are replaced by ONE flat view of that range."""
    groups = {}
    for t in tensors:
        groups.setdefault((t.untyped_storage().data_ptr(), t.dtype), []).append(t)
    out: List[torch.Tensor] = []
    for (_, _dt), ts in groups.items():
        ts = sorted(ts, key=lambda t: t.storage_offset())
        tiled = len(ts) > 1 and all(t.is_contiguous() for t in ts) and all(
            a.storage_offset() + a.numel() == b.storage_offset() for a, b in zip(ts, ts[1:]))
        if tiled:
            lo = ts[0].storage_offset()
            total = ts[-1].storage_offset() + ts[-1].numel() - lo
            out.append(ts[0].as_strided((total,), (1,), lo))
        else:
            out.extend(ts)
    return out


def begin_radii_reduce(radii: torch.Tensor, world: int, group=None):
    """``radii`` is final after the forward: start its MAX all-reduce then, underneath the backward.
    Pass the returned handle to :func:`allreduce_step`."""
    if world <= 1 or not dist.is_initialized():
        return None
    return dist.all_reduce(radii, op=dist.ReduceOp.MAX, group=group, async_op=True)


def allreduce_step(param_grads: Sequence[torch.Tensor], ndc_grad: Optional[torch.Tensor], radii: torch.Tensor,
                   world: int, group=None, average: bool = True, radii_work=None) -> torch.Tensor:
    """In-place exchange after a local backward.  Parameter gradients are averaged, ``ndc_grad``
    summed, ``radii`` max-reduced; returns the batch visibility (``max radii > 0`` -- no collective
    of its own is needed).  Works on NCCL and gloo.

    ``average=False``: the caller's loss already carries the 1/world factor (the reference's mean
    over the stacked batch, pointrix/model/loss.py:27-46, applies it to every view's loss and
    therefore to every view's ``ndc.grad`` too): everything is summed, no scaling pass, and the
    gradients of the fused backward -- views of one flat buffer -- go out as ONE all-reduce."""
    if world <= 1 or not dist.is_initialized():
        return radii > 0
    works = []
    grads = [g for g in param_grads if g is not None]
    if average:
        inv = 1.0 / world
        for g in grads:
            g.mul_(inv)
    bufs = grads + ([ndc_grad] if ndc_grad is not None else [])
    for t in _coalesce(bufs):
        works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
    if radii_work is None:
        radii_work = dist.all_reduce(radii, op=dist.ReduceOp.MAX, group=group, async_op=True)
    works.append(radii_work)
    for w in works:
        w.wait()
    return radii > 0
