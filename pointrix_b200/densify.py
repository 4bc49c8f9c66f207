"""Densification control of the Gaussian table -- the every-``duplicate_interval`` iterations clone / split / prune /
opacity-reset surgery of ``DensificationController`` (``pointrix/controller/gs.py:72-234, 286-313, 339-341``;
``BaseDensificationController.f_step / densify``, ``pointrix/controller/base.py:89-108``) on top of
:class:`pointrix_b200.optim.GaussianAdam` and :class:`~pointrix_b200.optim.DensificationStats`.

This is host logic, not a kernel: it runs once per hundred iterations, reshapes every tensor of the table and the
optimizer state (``pointrix/model/point_cloud/utils/point_utils.py:51-106``: moments of new points are zeros, moments of
replaced attributes are zeroed, the step counts stay), and is written with device-agnostic torch indexing so that it is
checked on CPU against the reference's own code executed where it lies (``tests/test_host_logic.py``).  The per-iteration
part of the controller -- the statistics -- is the fused kernel behind :meth:`GaussianAdam.step`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import Tensor

ADAM_STATES = ("exp_avg", "exp_avg_sq")  # point_utils.py:51


@dataclass
class DensifyConfig:
    """``BaseDensificationController.Config`` + ``DensificationController.Config`` (base.py:32-40, gs.py:39-46)."""
    prune_interval: int = 100
    min_opacity: float = 0.005
    duplicate_interval: int = 100
    densify_start_iter: int = 500
    densify_stop_iter: int = 15000
    densify_grad_threshold: float = 0.0002
    max_points: int = 5000000
    split_num: int = 2
    percent_dense: float = 0.01
    opacity_reset_interval: int = 3000
    normalize_grad: bool = True


# ---- optimizer-state surgery (point_utils.py:53-106), on plain dicts so that it runs anywhere -----------------------------
def extend_table(params: Dict[str, Tensor], state: Dict[str, Dict[str, Tensor]], new: Dict[str, Tensor]) -> None:
    """``extend_opt_by_tensor``: append rows; the new rows' moments are zeros."""
    for k, p in list(params.items()):
        ext = new[k].to(p.dtype)
        for key in ADAM_STATES:
            state[k][key] = torch.cat([state[k][key], torch.zeros_like(ext)], dim=0)
        params[k] = torch.cat([p.detach(), ext], dim=0).contiguous().requires_grad_(True)


def prune_table(params: Dict[str, Tensor], state: Dict[str, Dict[str, Tensor]], valid_mask: Tensor) -> None:
    """``reduce_opt_by_mask``: keep the rows where ``valid_mask`` is true."""
    for k, p in list(params.items()):
        for key in ADAM_STATES:
            state[k][key] = state[k][key][valid_mask]
        params[k] = p.detach()[valid_mask].contiguous().requires_grad_(True)


def replace_in_table(params: Dict[str, Tensor], state: Dict[str, Dict[str, Tensor]], new: Dict[str, Tensor]) -> None:
    """``replace_opt_tensor``: new values, zeroed moments."""
    for k, v in new.items():
        for key in ADAM_STATES:
            state[k][key] = torch.zeros_like(v)
        params[k] = v.detach().contiguous().requires_grad_(True)


def quat_to_rotmat(r: Tensor) -> Tensor:
    """``pointrix/utils/pose.py:84-117``: normalise (w first), then the rotation matrix of ``unitquat_to_rotmat``."""
    q = r / torch.sqrt((r * r).sum(dim=1))[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    m = torch.empty(q.shape[0], 3, 3, dtype=q.dtype, device=q.device)
    m[:, 0, 0] = x * x - y * y - z * z + w * w
    m[:, 1, 0] = 2 * (x * y + z * w)
    m[:, 2, 0] = 2 * (x * z - y * w)
    m[:, 0, 1] = 2 * (x * y - z * w)
    m[:, 1, 1] = -x * x + y * y - z * z + w * w
    m[:, 2, 1] = 2 * (y * z + x * w)
    m[:, 0, 2] = 2 * (x * z + y * w)
    m[:, 1, 2] = 2 * (y * z - x * w)
    m[:, 2, 2] = -x * x - y * y + z * z + w * w
    return m


def sigmoid_inv(x: Tensor) -> Tensor:
    return torch.log(x / (1 - x))  # point_utils.py:108


class DensificationController:
    """The reference controller's schedule and surgery.  ``optimizer``: anything with ``params`` (name -> leaf; RAW
    parameters: log-scales ``scaling``, logits ``opacity``, quaternions ``rotation``, ...) and ``state`` (name ->
    ``{"exp_avg", "exp_avg_sq"}``) dicts -- :class:`GaussianAdam`; ``stats``: a :class:`DensificationStats` (or any
    object with ``grad_accum / acc_steps / max_radii`` tensors and ``reset(n)`` / ``prune(mask)``)."""

    def __init__(self, optimizer, stats, cfg: Optional[DensifyConfig] = None, cameras_extent: float = 1.0):
        self.optimizer, self.stats = optimizer, stats
        self.cfg = cfg or DensifyConfig()
        self.cameras_extent = float(cameras_extent)
        self.step = 0

    # -- views of the table -------------------------------------------------------------------------------------------
    @property
    def params(self) -> Dict[str, Tensor]:
        return self.optimizer.params

    def __len__(self) -> int:
        return self.params["position"].shape[0]

    def _scaling(self) -> Tensor:
        return torch.exp(self.params["scaling"].detach())       # get_scaling, gaussian_points.py:74-76

    def _opacity(self) -> Tensor:
        return torch.sigmoid(self.params["opacity"].detach())   # get_opacity, gaussian_points.py:70-72

    # -- masks (gs.py:93-150) -----------------------------------------------------------------------------------------
    def generate_clone_mask(self, grads: Tensor) -> Tensor:
        mask = torch.norm(grads, dim=-1) >= self.cfg.densify_grad_threshold
        return torch.logical_and(mask, self._scaling().max(dim=1).values <= self.cfg.percent_dense * self.cameras_extent)

    def generate_split_mask(self, grads: Tensor) -> Tensor:
        padded = torch.zeros(len(self), device=grads.device, dtype=grads.dtype)
        padded[:grads.shape[0]] = grads.squeeze()   # points appended by the clone have no gradient yet
        mask = padded >= self.cfg.densify_grad_threshold
        return torch.logical_and(mask, self._scaling().max(dim=1).values > self.cfg.percent_dense * self.cameras_extent)

    def new_pos_scale(self, mask: Tensor, generator: Optional[torch.Generator] = None):
        """gs.py:152-180: ``split_num`` samples per selected Gaussian from N(0, diag(scaling^2)) rotated into the world
        frame; the new scales are the old ones / (0.8 * split_num), stored through the inverse activation (log)."""
        n = self.cfg.split_num
        scaling, rotation, position = self._scaling(), self.params["rotation"].detach(), self.params["position"].detach()
        stds = scaling[mask].repeat(n, 1)
        samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator)
        rots = quat_to_rotmat(rotation[mask]).repeat(n, 1, 1)
        new_pos = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + position[mask].repeat(n, 1)
        new_scaling = torch.log(scaling[mask].repeat(n, 1) / (0.8 * n))
        return new_pos, new_scaling

    # -- surgery (gs.py:182-234, 286-313, 72-91) ----------------------------------------------------------------------
    def _select(self, mask: Tensor) -> Dict[str, Tensor]:
        return {k: v.detach()[mask] for k, v in self.params.items()}

    def densify_clone(self, grads: Tensor) -> None:
        extend_table(self.params, self.optimizer.state, self._select(self.generate_clone_mask(grads)))
        self.stats.reset(len(self))

    def densify_split(self, grads: Tensor, generator: Optional[torch.Generator] = None) -> None:
        mask = self.generate_split_mask(grads)
        new_pos, new_scaling = self.new_pos_scale(mask, generator)
        new = self._select(mask)
        for k, v in new.items():
            new[k] = v.repeat(self.cfg.split_num, *([1] * (v.dim() - 1)))
        new["position"], new["scaling"] = new_pos, new_scaling
        extend_table(self.params, self.optimizer.state, new)
        self.stats.reset(len(self))
        valid = ~torch.cat((mask, torch.zeros(self.cfg.split_num * int(mask.sum()), device=mask.device, dtype=torch.bool)))
        prune_table(self.params, self.optimizer.state, valid)   # the split originals leave
        self.stats.prune(valid)

    def prune(self) -> None:
        size_threshold = 20 if self.step > self.cfg.opacity_reset_interval else None
        prune_filter = (self._opacity() < self.cfg.min_opacity).squeeze(-1)
        if size_threshold:
            prune_filter = prune_filter | (self.stats.max_radii > size_threshold)
            prune_filter = prune_filter | (self._scaling().max(dim=1).values > 0.1 * self.cameras_extent)
        valid = ~prune_filter
        prune_table(self.params, self.optimizer.state, valid)
        self.stats.prune(valid)

    def reset_opacity(self, reset_scale: float = 0.01) -> None:
        opc = self._opacity()
        replace_in_table(self.params, self.optimizer.state,
                         {"opacity": sigmoid_inv(torch.min(opc, torch.ones_like(opc) * reset_scale))})

    # -- schedule (base.py:89-108; gs.py:336-341) ---------------------------------------------------------------------
    def densify(self, generator: Optional[torch.Generator] = None) -> None:
        if self.step % self.cfg.duplicate_interval == 0:
            grads = self.stats.average_grad()
            self.densify_clone(grads)
            self.densify_split(grads, generator)
        if self.step % self.cfg.prune_interval == 0:
            self.prune()
        if self.step % self.cfg.opacity_reset_interval == 0:
            self.reset_opacity()

    def wants_statistics(self) -> bool:
        """Whether this iteration's statistics count (``f_step``, base.py:98-103: below ``max_points`` and before
        ``densify_stop_iter``): pass ``stats=`` to :meth:`GaussianAdam.step` only while this holds."""
        return len(self) < self.cfg.max_points and self.step < self.cfg.densify_stop_iter

    def f_step(self, generator: Optional[torch.Generator] = None) -> bool:
        """Call once per iteration AFTER the optimizer step that also updated the statistics
        (``GaussianAdam.step(stats=...)`` is the reference's ``preprocess``).  Returns True when the table changed
        (the caller's references to the parameter tensors are stale: re-read ``optimizer.params``)."""
        changed = False
        if len(self) < self.cfg.max_points and self.step < self.cfg.densify_stop_iter and self.step > self.cfg.densify_start_iter:
            due = (self.step % self.cfg.duplicate_interval == 0 or self.step % self.cfg.prune_interval == 0
                   or self.step % self.cfg.opacity_reset_interval == 0)
            if due:
                self.densify(generator)
                changed = True
        self.step += 1   # update_states, base.py:80-87
        return changed
