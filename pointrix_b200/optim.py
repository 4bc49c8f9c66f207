"""The step right after the render backward, fused: Adam on the Gaussian table and the
densification statistics, ONE kernel launch (``csrc/optim.cu`` behind ``pxb_adam_densify_step``).

Reference interfaces mirrored here
  * ``BaseOptimizer.update_model`` (``pointrix/optimizer/optimizer.py:128-140``): ``optimizer.step()`` then
    ``zero_grad(set_to_none=True)`` of a ``torch.optim.Adam`` with one parameter group per point-cloud attribute
    (names, learning rates and ``eps = 1e-15`` of ``examples/gaussian_splatting/configs/nerf.yaml:49-69``);
  * ``DensificationController.preprocess`` / ``accumulate_viewspace_grad`` / ``reset_controller_state``
    (``pointrix/controller/gs.py:250-333``): ``grad_accum``, ``acc_steps``, ``max_radii``.

There is no PyTorch fallback: CPU tensors are rejected like the render ops reject them.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Union

import torch
from torch import Tensor

from . import _lib
from .ops import _cuda, _p, _stream

# parameter groups and learning rates of examples/gaussian_splatting/configs/nerf.yaml:56-69
DEFAULT_LRS = {"position": 0.00016, "features": 0.0025, "features_rest": 0.000125, "scaling": 0.005, "rotation": 0.001,
               "opacity": 0.05}


class DensificationStats:
    """``grad_accum[P,1]``, ``acc_steps[P,1]``, ``max_radii[P]`` of ``DensificationController``
    (``reset_controller_state``, gs.py:250-257)."""

    def __init__(self, num_points: int, device, width: int, height: int, normalize_grad: bool = True):
        self.device = torch.device(device)
        self.width, self.height, self.normalize_grad = int(width), int(height), bool(normalize_grad)
        self.reset(num_points)

    def reset(self, num_points: int) -> None:
        self.grad_accum = torch.zeros((num_points, 1), device=self.device)
        self.acc_steps = torch.zeros((num_points, 1), device=self.device)
        self.max_radii = torch.zeros((num_points,), device=self.device)

    def prune(self, valid_points_mask: Tensor) -> None:
        """``prune_postprocess``, gs.py:236-247."""
        self.grad_accum = self.grad_accum[valid_points_mask]
        self.acc_steps = self.acc_steps[valid_points_mask]
        self.max_radii = self.max_radii[valid_points_mask]

    def average_grad(self) -> Tensor:
        """``grad_accum / acc_steps`` with NaN -> 0 (``BaseDensificationController.densify``, base.py:89-92)."""
        avg = self.grad_accum / self.acc_steps
        avg[avg.isnan()] = 0.0
        return avg

    def _scales(self):
        return (0.5 * self.width, 0.5 * self.height) if self.normalize_grad else (1.0, 1.0)

    def preprocess(self, uv_points: Union[Tensor, Sequence[Tensor]], visibility: Optional[Tensor], radii: Tensor) -> None:
        """``DensificationController.preprocess(uv_points=..., visibility=..., radii=...)``: the statistics alone
        (use :meth:`GaussianAdam.step` with ``stats=`` to fold them into the optimizer launch)."""
        _launch_step([], 0.9, 0.999, 1e-15, self, viewspace_grad(uv_points), radii)


def viewspace_grad(uv_points: Union[Tensor, Sequence[Tensor]]) -> Tensor:
    """Sum over the batch's views of ``ndc.grad`` (``accumulate_viewspace_grad``, gs.py:274-278).  ``uv_points``:
    the renderer's ``uv_points`` (one tensor or the list ``render_batch`` returns), or an already summed
    gradient tensor ``[P,2]`` that does not require grad (e.g. after the data-parallel exchange)."""
    if isinstance(uv_points, Tensor):
        g = uv_points.grad if uv_points.grad is not None else uv_points
        return g.reshape(-1, 2)
    grads = [vp.grad.reshape(-1, 2) for vp in uv_points]
    return grads[0] if len(grads) == 1 else torch.stack(grads, 0).sum(0)


def _launch_step(groups: List[_lib.AdamGroup], beta1, beta2, eps, stats: Optional[DensificationStats],
                 ndc_grad: Optional[Tensor], radii: Optional[Tensor]) -> None:
    arr = (_lib.AdamGroup * max(len(groups), 1))(*groups)
    if stats is not None:
        g = _cuda(ndc_grad, "uv_points.grad")
        if g.dtype != torch.float32 or not g.is_contiguous():
            g = g.float().contiguous()
        r = _cuda(radii, "radii").reshape(-1)
        if r.dtype != torch.int32 or not r.is_contiguous():
            r = r.to(torch.int32).contiguous()
        Pn = r.numel()
        if g.shape[0] != Pn or stats.max_radii.numel() != Pn:
            raise RuntimeError(f"densification statistics hold {stats.max_radii.numel()} points, got {g.shape[0]} gradients / {Pn} radii")
        sx, sy = stats._scales()
        dev = r.device
        tail = (Pn, _p(g), _p(r), sx, sy, _p(stats.grad_accum), _p(stats.acc_steps), _p(stats.max_radii))
        keep = (g, r)
    else:
        if not groups:
            return
        dev = torch.device("cuda", torch.cuda.current_device())
        tail = (0, _p(None), _p(None), 1.0, 1.0, _p(None), _p(None), _p(None))
        keep = ()
    with torch.cuda.device(dev):
        _lib.launch("pxb_adam_densify_step", C.cast(arr, C.c_void_p), len(groups), float(beta1), float(beta2), float(eps),
                    *tail, _stream(dev))
    del keep


class GaussianAdam:
    """Adam over the parameter groups of the Gaussian table -- ``torch.optim.Adam`` semantics (no weight decay, no
    amsgrad), one fused launch per step for all groups.

    ``params``: name -> leaf tensor (any of position / features / features_rest / scaling / rotation / opacity, or
    ``shs`` as ONE ``[P,16,3]`` leaf: its DC row then trains with ``lrs["features"]``, the rest with
    ``lrs["features_rest"]``, exactly as the reference's two groups do).  ``lrs``: name -> learning rate
    (mutable: a scheduler writes ``opt.lrs["position"] = ...``, as ``ExponLRScheduler`` does to the group's lr).
    ``state_dict()`` uses torch.optim.Adam's layout (``exp_avg``, ``exp_avg_sq``, ``step`` per parameter)."""

    def __init__(self, params: Dict[str, Tensor], lrs: Optional[Dict[str, float]] = None, betas=(0.9, 0.999),
                 eps: float = 1e-15):
        self.params = dict(params)
        self.lrs = dict(DEFAULT_LRS if lrs is None else lrs)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.steps: Dict[str, int] = {k: 0 for k in self.params}  # per parameter, as torch.optim keeps them
        self.state: Dict[str, Dict[str, Tensor]] = {}
        for k, p in self.params.items():
            _cuda(p, k)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"{k}: parameters must be contiguous fp32 tensors")
            need = ("features", "features_rest") if k == "shs" else (k,)
            for n in need:
                if n not in self.lrs:
                    raise KeyError(f"no learning rate for group {n!r}")
            self.state[k] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}

    # -- torch.optim.Optimizer-like surface ---------------------------------------------------------
    @property
    def param_groups(self) -> List[dict]:
        return [{"name": k, "params": [p], "lr": self.lrs.get(k, self.lrs.get("features")), "betas": self.betas,
                 "eps": self.eps} for k, p in self.params.items()]

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params.values():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def state_dict(self) -> dict:
        names = list(self.params)
        return {"state": {i: {"step": torch.tensor(float(self.steps[k])), "exp_avg": self.state[k]["exp_avg"],
                              "exp_avg_sq": self.state[k]["exp_avg_sq"]} for i, k in enumerate(names)},
                "param_groups": [{"name": k, "lr": g["lr"], "betas": self.betas, "eps": self.eps, "params": [i]}
                                 for i, (k, g) in enumerate(zip(names, self.param_groups))]}

    def load_state_dict(self, sd: dict) -> None:
        names = list(self.params)
        for i, k in enumerate(names):
            st = sd["state"][i]
            self.state[k]["exp_avg"].copy_(st["exp_avg"])
            self.state[k]["exp_avg_sq"].copy_(st["exp_avg_sq"])
            self.steps[k] = int(float(st["step"]))

    # -- the fused step -----------------------------------------------------------------------------
    def _groups(self, grads: Optional[Dict[str, Tensor]], keep: list) -> List[_lib.AdamGroup]:
        out = []
        for k, p in self.params.items():
            g = grads.get(k) if grads is not None else p.grad
            if g is None:
                continue  # torch.optim skips parameters without a gradient
            if g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
                g = g.float().contiguous().reshape(p.shape)
            keep.append(g)
            self.steps[k] += 1
            t = self.steps[k]
            st = self.state[k]
            rows = p.shape[0] if p.dim() > 0 else 1
            width = p.numel() // max(rows, 1)
            ptrs = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
            if k == "shs":
                # one [P,16,3] leaf, the reference's two groups: the DC row (columns 0..2) trains with the features
                # learning rate, the higher orders (3..47) with features_rest's
                out.append(_lib.AdamGroup(*ptrs, rows, 3, width, 0, width, 0, t, self.lrs["features"]))
                if width > 3:
                    out.append(_lib.AdamGroup(*ptrs, rows, width - 3, width, 3, width, 3, t, self.lrs["features_rest"]))
            else:
                out.append(_lib.AdamGroup(*ptrs, rows, width, width, 0, width, 0, t, self.lrs[k]))
        return out

    def step(self, stats: Optional[DensificationStats] = None, uv_points=None, visibility: Optional[Tensor] = None,
             radii: Optional[Tensor] = None, grads: Optional[Dict[str, Tensor]] = None) -> None:
        """One Adam step on every group that has a gradient (``.grad``, or ``grads[name]``).  With ``stats`` the
        densification statistics of this iteration (``uv_points``, ``radii`` as the renderer returned them;
        ``visibility`` is ``radii > 0`` and is recomputed in the kernel) are updated by the same launch."""
        keep: list = []
        groups = self._groups(grads, keep)
        _launch_step(groups, self.betas[0], self.betas[1], self.eps, stats,
                     viewspace_grad(uv_points) if stats is not None else None, radii)

    def update_model(self, **kwargs) -> None:
        """``BaseOptimizer.update_model`` (optimizer.py:128-140): step, then ``zero_grad(set_to_none=True)``."""
        with torch.no_grad():
            self.step(**{k: v for k, v in kwargs.items() if k in ("stats", "uv_points", "visibility", "radii", "grads")})
            self.zero_grad(set_to_none=True)
