"""Photometric losses at the render boundary, same names and argument meaning as
``pointrix/model/loss.py`` (``l1_loss`` :27-46, ``l2_loss`` :48-67, ``psnr`` :10-25, ``gaussian`` :69-71,
``create_window`` :119-123, ``ssim`` :73-97) plus the fused training loss ``BaseModel.get_loss_dict``
builds from them (``pointrix/model/base_model.py:96-124``).

Every loss is ONE hand-written CUDA pass forward and one backward (``csrc/loss.cu``) behind the C ABI;
torch holds the memory, the stream and the autograd edge.  There is no PyTorch fallback: CPU tensors
are rejected like the render ops reject them.  Gradients flow into the first argument (the rendered
image); the ground truth is treated as a constant, which is how the reference uses it.
"""
from __future__ import annotations

from math import exp
from typing import Dict, Optional

import torch
from torch import Tensor

from ._lib import launch, lib
from .ops import _f32, _p, _stream

WINDOW_SIZE = 11  # the only window the kernels implement (the reference never passes another)


def _ws(dev, nbytes: int) -> Tensor:
    return torch.empty(max(int(nbytes), 4), dtype=torch.uint8, device=dev)


def _as4d(t: Tensor) -> Tensor:
    # F.conv2d accepts [C,H,W] as an unbatched image and the reference reads channels at dim -3
    if t.dim() == 3:
        return t.unsqueeze(0)
    if t.dim() != 4:
        raise RuntimeError(f"expected a [B,C,H,W] or [C,H,W] image, got {tuple(t.shape)}")
    return t


def _no_gt_grad(gt: Tensor) -> None:
    if gt.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("pointrix_b200.loss: the ground truth (second argument) is a constant; "
                                  "gradients are produced for the prediction only")


# ---------------------------------------------------------------------------
# l1 / l2                  pointrix/model/loss.py:27-67
# ---------------------------------------------------------------------------
class _PixelLoss(torch.autograd.Function):
    """mode 1: |pred-gt|, mode 2: (pred-gt)^2; returns (means [B], map or empty) with B = the first
    dimension (per-image means, psnr) or 1 (one global mean, l1_loss / l2_loss)."""

    @staticmethod
    def forward(ctx, pred, gt, mode, want_map, per_image):
        p, g = _f32(pred, "pred"), _f32(gt, "gt")
        dev = p.device
        B = p.shape[0] if (per_image and p.dim() > 1) else 1
        if B > 65535:
            raise RuntimeError(f"per-image losses support up to 65535 images, got {B}")
        n = p.numel() // B
        means = torch.empty(B, dtype=torch.float32, device=dev)
        vmap = torch.empty_like(p) if want_map else None
        if p.numel() == 0:
            return means.fill_(float("nan")), (vmap if want_map else torch.empty(0, device=dev))
        ws = _ws(dev, 4 * B * min((n + 255) // 256, 1184))
        with torch.cuda.device(dev):
            launch("pxb_pixel_loss_forward", int(mode), B, n, _p(p), _p(g), _p(vmap), _p(means), _p(ws), ws.numel(),
                   _stream(dev))
        ctx.save_for_backward(p, g)
        ctx.mode, ctx.B, ctx.n, ctx.shape = int(mode), B, n, pred.shape
        return means, (vmap if want_map else torch.empty(0, device=dev))

    @staticmethod
    def backward(ctx, g_means, g_map):
        p, g = ctx.saved_tensors
        dev = p.device
        w = None if g_means is None else (g_means.float() / ctx.n).contiguous()
        gm = None
        if g_map is not None and g_map.numel() == p.numel():
            gm = g_map.float().contiguous()
        d = torch.empty_like(p)
        with torch.cuda.device(dev):
            launch("pxb_pixel_loss_backward", ctx.mode, ctx.B, ctx.n, _p(p), _p(g), _p(w), _p(gm), _p(d), _stream(dev))
        return d.view(ctx.shape), None, None, None, None


def _pixel_loss(pred: Tensor, gt: Tensor, mode: int, return_mean: bool) -> Tensor:
    assert pred.shape == gt.shape, "The shape of the two tensor should be the same."
    _no_gt_grad(gt)
    means, vmap = _PixelLoss.apply(pred, gt, mode, not return_mean, False)
    if not return_mean:
        return vmap.view(pred.shape)
    return means.reshape(())


def l1_loss(pred: Tensor, gt: Tensor, return_mean: bool = True) -> Tensor:
    """``torch.abs(pred - gt)`` (``.mean()`` unless ``return_mean=False``); loss.py:27-46."""
    return _pixel_loss(pred, gt, 1, return_mean)


def l2_loss(pred: Tensor, gt: Tensor, return_mean: bool = True) -> Tensor:
    """``(pred - gt) ** 2`` (``.mean()`` unless ``return_mean=False``); loss.py:48-67."""
    return _pixel_loss(pred, gt, 2, return_mean)


def psnr(img_pred: Tensor, img_gt: Tensor) -> Tensor:
    """Per-image PSNR ``[B,1]`` for images in [0,1]; loss.py:10-25."""
    assert img_pred.shape == img_gt.shape, "The shape of the two images should be the same."
    means, _ = _PixelLoss.apply(img_pred, img_gt, 2, False, True)
    return 20 * torch.log10(1.0 / torch.sqrt(means.view(-1, 1)))


# ---------------------------------------------------------------------------
# ssim                     pointrix/model/loss.py:69-123
# ---------------------------------------------------------------------------
def gaussian(window_size: int, sigma: float) -> Tensor:
    """Normalised 1-D Gaussian window, bit-identical to the reference's (loss.py:69-71): Python-double taps
    rounded to fp32, then an fp32 division by their fp32 sum.  Host helper."""
    centre = window_size // 2
    taps = torch.tensor([exp(-((k - centre) ** 2) / float(2 * sigma ** 2)) for k in range(window_size)],
                        dtype=torch.float32)
    return taps / taps.sum()


def create_window(window_size: int, channel: int) -> Tensor:
    """The reference's [channel,1,ws,ws] depthwise window (loss.py:119-123): outer product of the 1-D window
    (sigma 1.5) with itself.  Host helper; the kernels apply the 1-D window separably."""
    w = gaussian(window_size, 1.5)
    return torch.outer(w, w).expand(channel, 1, window_size, window_size).contiguous()


class _L1Ssim(torch.autograd.Function):
    """One pass: per-image mean |pred-gt| and per-image mean SSIM, [B] each."""

    @staticmethod
    def forward(ctx, pred, gt):
        p, g = _as4d(_f32(pred, "img1")), _as4d(_f32(gt, "img2"))
        dev = p.device
        B, Cc, H, W = p.shape
        l1 = torch.empty(B, dtype=torch.float32, device=dev)
        ss = torch.empty(B, dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[0]
        dmaps = torch.empty((3, B, Cc, H, W), dtype=torch.float32, device=dev) if need else None
        ws = _ws(dev, lib.pxb_loss_workspace_bytes(B, Cc, H, W))
        with torch.cuda.device(dev):
            launch("pxb_l1_ssim_forward", B, Cc, H, W, _p(p), _p(g), _p(dmaps), _p(l1), _p(ss), _p(ws), ws.numel(),
                   _stream(dev))
        if need:
            ctx.save_for_backward(p, g, dmaps)
        ctx.dims, ctx.shape = (B, Cc, H, W), pred.shape
        return l1, ss

    @staticmethod
    def backward(ctx, g_l1, g_ss):
        p, g, dmaps = ctx.saved_tensors
        B, Cc, H, W = ctx.dims
        inv_n = 1.0 / (Cc * H * W)
        g_l1 = None if g_l1 is None else g_l1.float().contiguous()
        g_ss = None if g_ss is None else g_ss.float().contiguous()
        d = torch.empty_like(p)
        with torch.cuda.device(p.device):
            launch("pxb_l1_ssim_backward", B, Cc, H, W, _p(p), _p(g), _p(dmaps), _p(g_l1), _p(g_ss), 1, inv_n, inv_n,
                   _p(d), _stream(p.device))
        return d.view(ctx.shape), None


class _L1SsimLoss(torch.autograd.Function):
    """The whole training loss in one pass: returns (loss, L1, 1-SSIM) as 0-dim tensors."""

    @staticmethod
    def forward(ctx, pred, gt, lambda_ssim):
        p, g = _as4d(_f32(pred, "img1")), _as4d(_f32(gt, "img2"))
        dev = p.device
        B, Cc, H, W = p.shape
        out3 = torch.empty(3, dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[0]
        dmaps = torch.empty((3, B, Cc, H, W), dtype=torch.float32, device=dev) if need else None
        ws = _ws(dev, lib.pxb_loss_workspace_bytes(B, Cc, H, W))
        with torch.cuda.device(dev):
            launch("pxb_l1_ssim_loss_forward", B, Cc, H, W, _p(p), _p(g), float(lambda_ssim), _p(dmaps), _p(out3),
                   _p(ws), ws.numel(), _stream(dev))
        if need:
            ctx.save_for_backward(p, g, dmaps)
        ctx.dims, ctx.shape, ctx.lam = (B, Cc, H, W), pred.shape, float(lambda_ssim)
        ctx.set_materialize_grads(False)
        return out3[0], out3[1], out3[2]

    @staticmethod
    def backward(ctx, g_loss, g_l1, g_sl):
        p, g, dmaps = ctx.saved_tensors
        B, Cc, H, W = ctx.dims
        inv_n, lam = 1.0 / (B * Cc * H * W), ctx.lam
        d = torch.empty_like(p)
        if g_l1 is None and g_sl is None and g_loss is not None:
            # the training case: only `loss` is differentiated; its upstream scalar stays on the device
            gl = g_loss.float().contiguous()
            args = (_p(gl), _p(gl), 0, (1.0 - lam) * inv_n, -lam * inv_n)
        else:
            z = torch.zeros((), dtype=torch.float32, device=p.device)
            gl, g1, gs = (z if t is None else t.float() for t in (g_loss, g_l1, g_sl))
            w_l1 = ((1.0 - lam) * gl + g1).contiguous()
            w_ss = (lam * gl + gs).contiguous()
            args = (_p(w_l1), _p(w_ss), 0, inv_n, -inv_n)
        with torch.cuda.device(p.device):
            launch("pxb_l1_ssim_backward", B, Cc, H, W, _p(p), _p(g), _p(dmaps), *args, _p(d), _stream(p.device))
        return d.view(ctx.shape), None, None


def _check_images(img1: Tensor, img2: Tensor, window_size: int) -> None:
    assert img1.shape == img2.shape, "The shape of the two images should be the same."
    if window_size != WINDOW_SIZE:
        raise ValueError(f"pointrix_b200.loss.ssim implements window_size={WINDOW_SIZE} only (got {window_size})")
    if img1.numel() == 0:
        raise RuntimeError("ssim of an empty image")
    _no_gt_grad(img2)


def ssim(img1: Tensor, img2: Tensor, window_size: int = 11, size_average: bool = True) -> Tensor:
    """SSIM with an 11x11 Gaussian window (sigma 1.5, zero padding): the mean over everything, or per
    image ``[B]`` with ``size_average=False``; loss.py:73-117."""
    _check_images(img1, img2, window_size)
    _, ss = _L1Ssim.apply(img1, img2)
    return ss.mean() if size_average else ss


def l1_ssim_loss(pred: Tensor, gt: Tensor, lambda_ssim: float = 0.2) -> Dict[str, Tensor]:
    """``{"loss", "L1_loss", "ssim_loss"}`` of ``BaseModel.get_loss_dict`` (base_model.py:113-124):
    ``loss = (1-lambda)*L1 + lambda*(1-SSIM)``, both terms from one kernel pass."""
    _check_images(pred, gt, WINDOW_SIZE)
    loss, L1_loss, ssim_loss = _L1SsimLoss.apply(pred, gt, float(lambda_ssim))
    return {"loss": loss, "L1_loss": L1_loss, "ssim_loss": ssim_loss}


def get_loss_dict(render_results: Dict[str, Tensor], batch, lambda_ssim: float = 0.2,
                  gt_images: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Function form of ``BaseModel.get_loss_dict(render_results, batch, step)``: stacks the batch's
    ``"image"`` entries (base_model.py:113-116) unless ``gt_images`` is given already stacked."""
    if gt_images is None:
        gt_images = torch.stack([batch[i]["image"] for i in range(len(batch))], dim=0)
    return l1_ssim_loss(render_results["rgb"], gt_images, lambda_ssim)
