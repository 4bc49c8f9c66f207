"""CPU restatement of the optimizer step and the densification statistics -- TEST INFRASTRUCTURE ONLY.

Only tests/ and bench.py's baseline legs may import this module; the product (pointrix_b200/) never does.
* Adam: the reference's optimizer IS ``torch.optim.Adam`` (``BaseOptimizer`` wraps it, pointrix/optimizer/
  optimizer.py:107-140; groups / learning rates / eps of examples/gaussian_splatting/configs/nerf.yaml:49-69),
  so the oracle is torch.optim.Adam itself, run on CPU tensors.
* Statistics: ``DensificationController.accumulate_viewspace_grad`` + ``preprocess`` (pointrix/controller/
  gs.py:259-284, 316-333) restated below.  Pinned: tests/golden/ref_controller.npz holds inputs and outputs of
  those two methods of the reference's own gs.py executed where it lies
  (``python oracle/make_golden.py --from-ref-controller``); tests/test_oracle.py checks this file against them.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

NERF_YAML_LRS = {"position": 0.00016, "features": 0.0025, "features_rest": 0.000125, "scaling": 0.005,
                 "rotation": 0.001, "opacity": 0.05}  # nerf.yaml:56-69


def make_adam(params: Dict[str, torch.Tensor], lrs: Dict[str, float] = NERF_YAML_LRS, eps: float = 1e-15):
    """torch.optim.Adam with one group per attribute, as the reference's parser builds it
    (pointrix/optimizer/__init__.py; nerf.yaml:49-69)."""
    return torch.optim.Adam([{"params": [p], "lr": lrs[k], "name": k} for k, p in params.items()], eps=eps)


def accumulate_viewspace_grad(uv_grads: Sequence[torch.Tensor], width: int, height: int, normalize_grad: bool = True):
    # gs.py:259-284: sum of every view's ndc.grad, then x *= W/2, y *= H/2
    g = torch.zeros_like(uv_grads[0])
    for vg in uv_grads:
        g += vg.clone()
    if normalize_grad:
        g[..., 0] *= width / 2.0
        g[..., 1] *= height / 2.0
    return g


def densify_preprocess(grad_accum, acc_steps, max_radii, uv_grads: List[torch.Tensor], visibility, radii, width, height,
                       normalize_grad: bool = True) -> None:
    """In place, gs.py:316-333: max_radii[sel] = max(max_radii[sel], radii[sel]);
    grad_accum[sel] += ||viewspace_grad[sel, :2]||; acc_steps[sel] += 1."""
    point_grad = accumulate_viewspace_grad(uv_grads, width, height, normalize_grad)
    sel = visibility
    max_radii[sel] = torch.max(max_radii[sel], radii[sel])
    grad_accum[sel] += torch.norm(point_grad[sel, :2], dim=-1, keepdim=True)
    acc_steps[sel] += 1
