"""msplat-compatible operator API on top of the C ABI (no CPU fallback).

Same names, argument meaning and error behaviour as the reference's Python op
layer (``msplat/msplat/*.py``): ``project_point``, ``compute_cov3d``,
``ewa_project``, ``sort_gaussian``, ``compute_sh``, ``alpha_blending`` and the
one-line ``rasterization`` pipeline (``msplat/msplat/__init__.py:22-93``).
Each differentiable op is a ``torch.autograd.Function`` whose forward/backward
call one C entry point on PyTorch's *current* stream; torch is used only for
device memory, streams and autograd plumbing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import launch, lib

TILE = 16
MAX_CH = 26  # PXB_MAX_CHANNELS_PER_PASS


def _cuda(t: Tensor, name: str) -> Tensor:
    # the reference's CHECK_INPUT (msplat/msplat/include/utils.h:9-10)
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    return t


def _f32(t: Tensor, name: str) -> Tensor:
    t = _cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32(t: Tensor, name: str) -> Tensor:
    t = _cuda(t, name)
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def _vis(visible: Optional[Tensor], name: str = "visible") -> Optional[Tensor]:
    if visible is None:
        return None
    v = _cuda(visible, name).reshape(-1)
    if v.dtype != torch.bool:
        v = v != 0
    return v.contiguous()


def _p(t: Optional[Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev) -> C.c_void_p:
    _lib.ensure_init()
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _aligned16(t: Tensor) -> Tensor:
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


# ---------------------------------------------------------------------------
# project_point          msplat/msplat/project_point.py:8-98
# ---------------------------------------------------------------------------
class _ProjectPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, intr, extr, W, H, nearest, extent):
        xyz_c, intr_c, extr_c = _f32(xyz, "xyz"), _f32(intr, "intr"), _f32(extr, "extr")
        P = xyz_c.shape[0]
        uv = torch.empty(P, 2, dtype=torch.float32, device=xyz_c.device)
        depth = torch.empty(P, 1, dtype=torch.float32, device=xyz_c.device)
        with torch.cuda.device(xyz_c.device):
            launch("pxb_project_point_forward", P, _p(xyz_c), _p(intr_c), _p(extr_c), int(W), int(H), float(nearest),
                                                float(extent), _p(uv), _p(depth), _stream(xyz_c.device))
        ctx.save_for_backward(xyz_c, intr_c, extr_c, depth)
        ctx.shapes = (intr.shape, extr.shape)
        return uv, depth

    @staticmethod
    def backward(ctx, dL_duv, dL_ddepth):
        xyz, intr, extr, depth = ctx.saved_tensors
        P = xyz.shape[0]
        dev = xyz.device
        dL_duv = _f32(dL_duv, "dL_duv")
        dL_ddepth = _f32(dL_ddepth, "dL_ddepth")
        d_xyz = torch.empty_like(xyz)
        need_i, need_e = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_intr = torch.zeros(4, dtype=torch.float32, device=dev) if need_i else None
        d_extr = torch.zeros(12, dtype=torch.float32, device=dev) if need_e else None
        with torch.cuda.device(dev):
            launch("pxb_project_point_backward", P, _p(xyz), _p(intr), _p(extr), _p(depth), _p(dL_duv), _p(dL_ddepth),
                                                 _p(d_xyz), _p(d_intr), _p(d_extr), _stream(dev))
        gi = d_intr.reshape(ctx.shapes[0]) if need_i else None
        ge = d_extr.reshape(ctx.shapes[1]) if need_e else None
        return d_xyz, gi, ge, None, None, None, None


def project_point(xyz: Tensor, intr: Tensor, extr: Tensor, W: int, H: int, nearest: float = 0.0,
                  extent: float = 1.3) -> Tuple[Tensor, Tensor]:
    """Project 3D points to the screen -> (uv[P,2], depth[P,1]); culled points are zeros."""
    return _ProjectPoint.apply(xyz, intr, extr, W, H, nearest, extent)


# ---------------------------------------------------------------------------
# compute_cov3d          msplat/msplat/compute_cov3d.py:7-64
# ---------------------------------------------------------------------------
class _ComputeCov3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scales, uquats, visible):
        s, q = _f32(scales, "scales"), _aligned16(_f32(uquats, "uquats"))
        v = _vis(visible)
        P = s.shape[0]
        cov = torch.empty(P, 6, dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            launch("pxb_compute_cov3d_forward", P, _p(s), _p(q), _p(v), _p(cov), _stream(s.device))
        ctx.save_for_backward(s, q, v if v is not None else torch.empty(0, device=s.device))
        ctx.has_vis = v is not None
        return cov

    @staticmethod
    def backward(ctx, dL_dcov3d):
        s, q, v = ctx.saved_tensors
        v = v if ctx.has_vis else None
        P = s.shape[0]
        g = _f32(dL_dcov3d, "dL_dcov3d")
        d_s = torch.empty_like(s)
        d_q = torch.empty_like(q)
        with torch.cuda.device(s.device):
            launch("pxb_compute_cov3d_backward", P, _p(s), _p(q), _p(v), _p(g), _p(d_s), _p(d_q), _stream(s.device))
        return d_s, d_q, None


def compute_cov3d(scales: Tensor, uquats: Tensor, visible: Optional[Tensor] = None) -> Tensor:
    """3D covariance upper triangle [P,6] from scales and unit quaternions (w,x,y,z)."""
    return _ComputeCov3D.apply(scales, uquats, visible)


# ---------------------------------------------------------------------------
# ewa_project            msplat/msplat/ewa_project.py:8-94
# ---------------------------------------------------------------------------
class _EWAProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, cov3d, intr, extr, uv, W, H, visible):
        xyz_c, cov_c = _f32(xyz, "xyz"), _f32(cov3d, "cov3d")
        intr_c, extr_c, uv_c = _f32(intr, "intr"), _f32(extr, "extr"), _f32(uv, "uv")
        v = _vis(visible)
        P = xyz_c.shape[0]
        dev = xyz_c.device
        conic = torch.empty(P, 3, dtype=torch.float32, device=dev)
        radius = torch.empty(P, dtype=torch.int32, device=dev)
        tiles = torch.empty(P, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            launch("pxb_ewa_project_forward", P, _p(xyz_c), _p(cov_c), _p(intr_c), _p(extr_c), _p(uv_c), int(W), int(H),
                                              _p(v), _p(conic), _p(radius), _p(tiles), _stream(dev))
        ctx.save_for_backward(xyz_c, cov_c, intr_c, extr_c, radius)
        ctx.shapes = (intr.shape, extr.shape)
        ctx.mark_non_differentiable(radius, tiles)
        return conic, radius, tiles

    @staticmethod
    def backward(ctx, dL_dconic, _dr, _dt):
        xyz, cov3d, intr, extr, radius = ctx.saved_tensors
        P = xyz.shape[0]
        dev = xyz.device
        g = _f32(dL_dconic, "dL_dconic")
        d_xyz = torch.empty_like(xyz)
        d_cov = torch.empty_like(cov3d)
        need_i, need_e = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        d_intr = torch.zeros(4, dtype=torch.float32, device=dev) if need_i else None
        d_extr = torch.zeros(12, dtype=torch.float32, device=dev) if need_e else None
        with torch.cuda.device(dev):
            launch("pxb_ewa_project_backward", P, _p(xyz), _p(cov3d), _p(intr), _p(extr), _p(radius), _p(g), _p(d_xyz),
                                               _p(d_cov), _p(d_intr), _p(d_extr), _stream(dev))
        gi = d_intr.reshape(ctx.shapes[0]) if need_i else None
        ge = d_extr.reshape(ctx.shapes[1]) if need_e else None
        return d_xyz, d_cov, gi, ge, None, None, None, None


def ewa_project(xyz: Tensor, cov3d: Tensor, intr: Tensor, extr: Tensor, uv: Tensor, W: int, H: int,
                visible: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """EWA projection -> (conic[P,3], radius[P] int32, tiles[P] int32)."""
    return _EWAProject.apply(xyz, cov3d, intr, extr, uv, W, H, visible)


# ---------------------------------------------------------------------------
# sort_gaussian          msplat/msplat/sort_gaussian.py:8-54
# ---------------------------------------------------------------------------
def _prepare(depth: Tensor, radius: Tensor, tiles: Tensor, stream):
    """depth-order the Gaussians + count intersections -> (total_dev[1], ws_p, ws_p_bytes)"""
    dev = depth.device
    P = radius.numel()
    total = torch.empty(1, dtype=torch.int32, device=dev)
    ws_p_bytes = lib.pxb_bin_prepare_workspace_bytes(P)
    ws_p = torch.empty(ws_p_bytes, dtype=torch.uint8, device=dev)
    launch("pxb_bin_prepare", P, _p(depth), _p(radius), _p(tiles), _p(total), _p(ws_p), ws_p_bytes, stream)
    return total, ws_p, ws_p_bytes


def _bin(uv_like: Tensor, uv_stride: int, depth: Tensor, radius: Tensor, tiles: Tensor, W: int, H: int,
         return_keys: bool = False, tight: bool = False):
    dev = depth.device
    P = radius.numel()
    n_tiles = ((W + TILE - 1) // TILE) * ((H + TILE - 1) // TILE)
    tile_range = torch.empty(n_tiles, 2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        stream = _stream(dev)
        total, ws_p, ws_p_bytes = _prepare(depth, radius, tiles, stream)
        N = int(total.item())  # the one host sync of the operator path (the reference has two)
        LAST_N[(dev.index, bool(tight))] = N
        idx_sorted = torch.empty(N, dtype=torch.int32, device=dev)
        keys = torch.empty(N, dtype=torch.int64, device=dev) if return_keys else None
        ws_n_bytes = lib.pxb_bin_sort_workspace_bytes(max(N, 1), int(W), int(H))
        ws_n = torch.empty(ws_n_bytes, dtype=torch.uint8, device=dev)
        launch("pxb_sort_gaussian", P, N, _p(None), _p(uv_like), uv_stride, int(tight), _p(depth), _p(radius), _p(tiles), int(W), int(H),
               _p(idx_sorted), _p(tile_range), _p(keys), _p(ws_p), ws_p_bytes, _p(ws_n), ws_n_bytes, stream)
    if return_keys:
        return idx_sorted, tile_range, keys
    return idx_sorted, tile_range


LAST_N = {}  # (device index, tight) -> intersection count of the most recent binning (diagnostics / bench)


def sort_gaussian(uv: Tensor, depth: Tensor, W: int, H: int, radius: Tensor, tiles: Tensor, return_keys: bool = False):
    """Sort Gaussians by [tile|depth] -> (idx_sorted[N] int32, tile_range[tiles,2] int32)."""
    uv_c, d_c = _f32(uv.detach(), "uv"), _f32(depth.detach(), "depth").reshape(-1)
    r_c, t_c = _i32(radius, "radius").reshape(-1), _i32(tiles, "tiles").reshape(-1)
    return _bin(uv_c, 2, d_c, r_c, t_c, W, H, return_keys)


# ---------------------------------------------------------------------------
# compute_sh             msplat/msplat/compute_sh.py:8-64
# ---------------------------------------------------------------------------
class _ComputeSH(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shs, view_dirs, visible):
        s, d = _f32(shs, "shs"), _f32(view_dirs, "view_dirs")
        v = _vis(visible)
        P, Cc, D = s.shape
        val = torch.empty(P, Cc, dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            launch("pxb_compute_sh_forward", P, Cc, D, _p(s), _p(d), _p(v), _p(val), _stream(s.device))
        ctx.save_for_backward(s, d, v if v is not None else torch.empty(0, device=s.device))
        ctx.has_vis = v is not None
        return val

    @staticmethod
    def backward(ctx, dL_dvalue):
        s, d, v = ctx.saved_tensors
        v = v if ctx.has_vis else None
        P, Cc, D = s.shape
        g = _f32(dL_dvalue, "dL_dvalue")
        d_s = torch.empty_like(s)
        d_d = torch.empty_like(d)
        with torch.cuda.device(s.device):
            launch("pxb_compute_sh_backward", P, Cc, D, _p(s), _p(d), _p(v), _p(g), _p(d_s), _p(d_d), _stream(s.device))
        return d_s, d_d, None


def compute_sh(shs: Tensor, view_dirs: Tensor, visible: Optional[Tensor] = None) -> Tensor:
    """value[P,C] = sum_d SH_d(dir) * shs[P,C,d]  (no +0.5, no clamp)."""
    return _ComputeSH.apply(shs, view_dirs, visible)


# ---------------------------------------------------------------------------
# alpha_blending         msplat/msplat/alpha_blending.py:7-135
# ---------------------------------------------------------------------------
def _chunks(Cc: int):
    c0 = 0
    while c0 < Cc:
        cn = min(MAX_CH, Cc - c0)
        yield c0, cn, lib.pxb_record_stride(cn)
        c0 += cn


class _AlphaBlending(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, ndc):
        uv_c, cn_c = _f32(uv, "uv"), _f32(conic, "conic")
        op_c, ft_c = _f32(opacity, "opacity"), _f32(feature, "feature")
        ids, tr = _i32(idx_sorted, "idx_sorted"), _i32(tile_range, "tile_range")
        P, Cc = ft_c.shape
        dev = ft_c.device
        W, H, bg = int(W), int(H), float(bg)
        out = torch.empty(Cc, H, W, dtype=torch.float32, device=dev)
        final_T = torch.empty(H, W, dtype=torch.float32, device=dev)
        ncontrib = torch.empty(H, W, dtype=torch.int32, device=dev)
        rec_keep = None
        with torch.cuda.device(dev):
            stream = _stream(dev)
            for c0, cn, S in _chunks(Cc):
                rec = torch.empty(max(P, 1), S, dtype=torch.float32, device=dev)
                launch("pxb_pack_records", P, _p(uv_c), _p(cn_c), _p(op_c), _p(ft_c), Cc, c0, cn, S, _p(rec), stream)
                launch("pxb_blend_forward", _p(rec), S, cn, _p(ids), _p(tr), bg, W, H, _p(final_T), _p(ncontrib),
                                            _p(out[c0:]), stream)
                if Cc <= MAX_CH:
                    rec_keep = rec
        ctx.W, ctx.H, ctx.bg = W, H, bg
        ctx.has_ndc = ndc is not None
        ctx.single = rec_keep is not None
        ctx.save_for_backward(uv_c, cn_c, op_c, ft_c, ids, tr, final_T, ncontrib,
                              rec_keep if rec_keep is not None else torch.empty(0, device=dev))
        return out

    @staticmethod
    def backward(ctx, dL_drendered):
        uv, conic, opacity, feature, ids, tr, final_T, ncontrib, rec_keep = ctx.saved_tensors
        W, H, bg = ctx.W, ctx.H, ctx.bg
        P, Cc = feature.shape
        dev = feature.device
        g = _f32(dL_drendered, "dL_drendered")
        d_uv = torch.empty(P, 2, dtype=torch.float32, device=dev)
        d_conic = torch.empty(P, 3, dtype=torch.float32, device=dev)
        d_op = torch.empty(opacity.shape, dtype=torch.float32, device=dev)
        d_feat = torch.empty(P, Cc, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = _stream(dev)
            for n, (c0, cn, S) in enumerate(_chunks(Cc)):
                if ctx.single:
                    rec = rec_keep
                else:
                    rec = torch.empty(max(P, 1), S, dtype=torch.float32, device=dev)
                    launch("pxb_pack_records", P, _p(uv), _p(conic), _p(opacity), _p(feature), Cc, c0, cn, S, _p(rec), stream)
                grec = torch.zeros(max(P, 1), S, dtype=torch.float32, device=dev)
                launch("pxb_blend_backward", _p(rec), S, cn, _p(ids), _p(tr), bg, W, H, _p(final_T), _p(ncontrib),
                                             _p(g[c0:]), _p(grec), stream)
                launch("pxb_unpack_grads", P, _p(grec), S, Cc, c0, cn, 1 if n > 0 else 0, _p(d_uv), _p(d_conic), _p(d_op),
                                           _p(d_feat), stream)
        d_ndc = None
        if ctx.has_ndc:
            # msplat/msplat/alpha_blending.py:106-110
            d_ndc = d_uv * torch.tensor([0.5 * W, 0.5 * H], dtype=d_uv.dtype, device=dev)[None, :]
        return d_uv, d_conic, d_op, d_feat, None, None, None, None, None, d_ndc


def alpha_blending(uv: Tensor, conic: Tensor, opacity: Tensor, feature: Tensor, idx_sorted: Tensor, title_bins: Tensor,
                   bg: float, W: int, H: int, ndc: Optional[Tensor] = None) -> Tensor:
    """Alpha-blend sorted 2D Gaussians tile by tile -> feature map [C,H,W]."""
    return _AlphaBlending.apply(uv, conic, opacity, feature, idx_sorted, title_bins, bg, W, H, ndc)


def alpha_blending_aux(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H):
    """Forward only, also returning (final_T[H,W], ncontrib[H,W]) -- the tensors the
    reference's ``_C.alpha_blending_forward`` returns (alpha_blending.cu:248-394)."""
    uv_c, cn_c = _f32(uv, "uv"), _f32(conic, "conic")
    op_c, ft_c = _f32(opacity, "opacity"), _f32(feature, "feature")
    ids, tr = _i32(idx_sorted, "idx_sorted"), _i32(tile_range, "tile_range")
    P, Cc = ft_c.shape
    dev = ft_c.device
    out = torch.empty(Cc, H, W, dtype=torch.float32, device=dev)
    final_T = torch.empty(H, W, dtype=torch.float32, device=dev)
    ncontrib = torch.empty(H, W, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        stream = _stream(dev)
        for c0, cn, S in _chunks(Cc):
            rec = torch.empty(max(P, 1), S, dtype=torch.float32, device=dev)
            launch("pxb_pack_records", P, _p(uv_c), _p(cn_c), _p(op_c), _p(ft_c), Cc, c0, cn, S, _p(rec), stream)
            launch("pxb_blend_forward", _p(rec), S, cn, _p(ids), _p(tr), float(bg), int(W), int(H), _p(final_T),
                                        _p(ncontrib), _p(out[c0:]), stream)
    return out, final_T, ncontrib


# ---------------------------------------------------------------------------
# rasterization          msplat/msplat/__init__.py:22-93
# ---------------------------------------------------------------------------
def rasterization(xyz: Tensor, scale: Tensor, rotate: Tensor, opacity: Tensor, feature: Tensor, intr: Tensor,
                  extr: Tensor, W: int, H: int, bg: float, ndc: Optional[Tensor] = None) -> Tensor:
    """Vanilla 3DGS rasterization pipeline: project -> cov3d -> ewa -> sort -> blend."""
    (uv, depth) = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    (conic, radius, tiles_touched) = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    (gaussian_ids_sorted, tile_range) = sort_gaussian(uv, depth, W, H, radius, tiles_touched)
    return alpha_blending(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, bg, W, H, ndc)
