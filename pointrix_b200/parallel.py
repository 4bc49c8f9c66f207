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


def allreduce_step(param_grads: Sequence[torch.Tensor], ndc_grad: Optional[torch.Tensor], radii: torch.Tensor,
                   world: int, group=None) -> torch.Tensor:
    """In-place exchange after a local backward.  Parameter gradients are averaged,
    ``ndc_grad`` summed, ``radii`` max-reduced; returns the batch visibility
    (``max radii > 0`` -- no collective of its own is needed).  Works on NCCL and gloo."""
    if world <= 1 or not dist.is_initialized():
        return radii > 0
    works = []
    inv = 1.0 / world
    for g in param_grads:
        if g is None:
            continue
        g.mul_(inv)
        works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True))
    if ndc_grad is not None:
        works.append(dist.all_reduce(ndc_grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
    works.append(dist.all_reduce(radii, op=dist.ReduceOp.MAX, group=group, async_op=True))
    for w in works:
        w.wait()
    return radii > 0
