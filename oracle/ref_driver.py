"""Drive the compiled, unmodified reference (``oracle/_ref/msplat_ref_C*.so``) through
the exact call sequence of its own Python layer -- TEST INFRASTRUCTURE ONLY.

The reference's Python wrappers (msplat/msplat/*.py) and plugin
(pointrix/model/renderer/msplat.py:94-158) are thin sequencing code; they are
restated here (not copied) so that the 12 raw ``_C`` entry points
(msplat/msplat/src/ext.cpp:14-25) can be driven on the GPU box, where
``/root/reference`` does not exist.  Used by the ``-m gpu`` tests as the
bit-exact checker, by ``oracle/make_golden.py`` and by ``bench.py`` (``ref_gpu``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .build_ref import load_ref

_C = None


def available() -> bool:
    global _C
    if _C is None:
        _C = load_ref()
    return _C is not None and torch.cuda.is_available()


def C():
    if not available():
        raise RuntimeError("oracle/_ref is not built (python oracle/build_ref.py) or no GPU")
    return _C


def sort_gaussian(uv, depth, W, H, radius, tiles, return_keys=False):
    """msplat/msplat/sort_gaussian.py:42-52"""
    c = C()
    cum = torch.cumsum(tiles, dim=0, dtype=torch.int32)
    key, idx = c.compute_gaussian_key(uv, depth, W, H, radius, cum)
    key_sorted, perm = torch.sort(key)
    idx_sorted = torch.gather(idx, 0, perm)
    tile_range = c.compute_tile_gaussian_range(W, H, cum, key_sorted)
    if return_keys:
        return idx_sorted, tile_range, key_sorted
    return idx_sorted, tile_range


def render_forward(height, width, extrinsic_matrix, intrinsic_params, camera_center, position, opacity, scaling,
                   rotation, shs, sh_degree=3, bg=1.0, render_depth=False, extra: Optional[torch.Tensor] = None,
                   nearest=0.2) -> Dict[str, torch.Tensor]:
    """Forward of pointrix/model/renderer/msplat.py:94-151 with every intermediate kept."""
    c = C()
    W, H = int(width), int(height)
    direction = position - camera_center.reshape(1, 3)
    direction = direction / direction.norm(dim=1, keepdim=True)
    sh_coeff = shs.permute(0, 2, 1)
    sh_mask = torch.zeros_like(sh_coeff)
    sh_mask[..., : (sh_degree + 1) ** 2] = 1.0
    sh_in = (sh_coeff * sh_mask).contiguous()
    vis_all = torch.ones_like(sh_in[:, 0, 0], dtype=torch.bool)
    sh_val = c.compute_sh_forward(sh_in, direction, vis_all)
    rgb = (sh_val + 0.5).clamp(min=0.0)
    extr = extrinsic_matrix[:3, :].contiguous()
    intr = intrinsic_params.reshape(-1).contiguous()
    uv, depth = c.project_point_forward(position, intr, extr, W, H, float(nearest), 1.3)
    visible = (depth != 0).reshape(-1)
    cov3d = c.compute_cov3d_forward(scaling, rotation, visible)
    conic, radius, tiles = c.ewa_project_forward(position, cov3d, intr, extr, uv, W, H, visible)
    idx_sorted, tile_range, keys = sort_gaussian(uv, depth, W, H, radius, tiles, return_keys=True)
    cols = [rgb] + ([depth] if render_depth else []) + ([extra] if extra is not None else [])
    feature = torch.cat(cols, dim=-1).contiguous()
    img, final_T, ncontrib = c.alpha_blending_forward(uv, conic, opacity, feature, idx_sorted, tile_range, float(bg), W, H)
    return dict(direction=direction, sh_in=sh_in, sh_val=sh_val, rgb=rgb, extr=extr, intr=intr, uv=uv, depth=depth,
                visible=visible, cov3d=cov3d, conic=conic, radius=radius, tiles=tiles, idx_sorted=idx_sorted,
                tile_range=tile_range, keys=keys, feature=feature, img=img, final_T=final_T, ncontrib=ncontrib)


def render_backward(f: Dict[str, torch.Tensor], dL_dimg, position, opacity, scaling, rotation, shs, camera_center,
                    sh_degree=3, bg=1.0, render_depth=False, n_extra=0, camera_grads=False) -> Dict[str, torch.Tensor]:
    """Manual chain of the reference backward entry points in autograd order
    (msplat/msplat/*.py backward methods + the torch ops of msplat.py:94-151)."""
    c = C()
    H, W = f["img"].shape[1:]
    d_uv, d_conic, d_op, d_feat = c.alpha_blending_backward(
        f["uv"], f["conic"], opacity, f["feature"], f["idx_sorted"], f["tile_range"], float(bg), W, H, f["final_T"],
        f["ncontrib"], dL_dimg.contiguous())
    d_ndc = d_uv * torch.tensor([0.5 * W, 0.5 * H], device=d_uv.device)[None]
    d_rgb = d_feat[:, :3]
    d_depth = d_feat[:, 3:4].contiguous() if render_depth else torch.zeros_like(f["depth"])
    d_extra = d_feat[:, 3 + int(render_depth):] if n_extra else None
    intr = f["intr"].clone().requires_grad_(camera_grads)
    extr = f["extr"].clone().requires_grad_(camera_grads)
    d_xyz_ewa, d_cov3d, d_intr_e, d_extr_e = c.ewa_project_backward(position, f["cov3d"], intr, extr, f["radius"], d_conic)
    d_scale, d_quat = c.compute_cov3d_backward(scaling, rotation, f["visible"], d_cov3d)
    d_xyz_proj, d_intr_p, d_extr_p = c.project_point_backward(position, intr, extr, W, H, f["uv"], f["depth"], d_uv, d_depth)
    d_val = (d_rgb * (f["sh_val"] + 0.5 > 0)).contiguous()
    vis_all = torch.ones_like(f["sh_in"][:, 0, 0], dtype=torch.bool)
    d_shin, d_dir = c.compute_sh_backward(f["sh_in"], f["direction"], vis_all, d_val)
    sh_mask = torch.zeros_like(d_shin)
    sh_mask[..., : (sh_degree + 1) ** 2] = 1.0
    d_shs = (d_shin * sh_mask).permute(0, 2, 1).contiguous()
    # normalize backward
    v = position - camera_center.reshape(1, 3)
    n = v.norm(dim=1, keepdim=True)
    d = v / n
    d_v = (d_dir - d * (d_dir * d).sum(1, keepdim=True)) / n
    out = dict(position=d_xyz_ewa + d_xyz_proj + d_v, scaling=d_scale, rotation=d_quat, opacity=d_op, shs=d_shs,
               ndc=d_ndc, d_uv=d_uv, d_conic=d_conic, d_feat=d_feat, camera_center=-d_v.sum(0))
    if d_extra is not None:
        out["extra"] = d_extra
    if camera_grads:
        out["intr"] = d_intr_e + d_intr_p
        out["extr"] = d_extr_e + d_extr_p
    return out
