"""CPU restatement of the reference's photometric loss -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module (bench.py times
it as "the reference's torch-op formulation" beside the fused kernels, never as part of them; tools/ are
developer scripts); the product (pointrix_b200/) never does.  Plain PyTorch on CPU tensors, each function citing the reference lines
it follows.  Pinned: tests/golden/ref_loss.npz holds inputs, values and autograd gradients produced by
the reference's own pointrix/model/loss.py imported where it lies (generator:
`python oracle/make_golden.py --from-ref-loss`), and tests/test_oracle.py checks this file against them.
"""
from __future__ import annotations

from math import exp

import torch
import torch.nn.functional as F


def l1_loss(pred, gt, return_mean=True):
    # pointrix/model/loss.py:27-46
    assert pred.shape == gt.shape
    d = torch.abs(pred - gt)
    return d.mean() if return_mean else d


def l2_loss(pred, gt, return_mean=True):
    # pointrix/model/loss.py:48-67
    assert pred.shape == gt.shape
    d = (pred - gt) ** 2
    return d.mean() if return_mean else d


def psnr(img_pred, img_gt):
    # pointrix/model/loss.py:10-25
    l2 = l2_loss(img_pred, img_gt, return_mean=False)
    m = l2.reshape(img_pred.shape[0], -1).mean(1, keepdim=True)
    return 20 * torch.log10(1.0 / torch.sqrt(m))


def window_1d(window_size=11, sigma=1.5):
    """The reference's 1-D window (gaussian(), pointrix/model/loss.py:69-71): each tap is a Python-double
    exp rounded to fp32, the normalisation an fp32 division by the fp32 sum."""
    half = window_size // 2
    taps = [exp(-((i - half) ** 2) / float(2 * sigma ** 2)) for i in range(window_size)]
    w = torch.tensor(taps, dtype=torch.float32)
    return w / w.sum()


def _box(x, kernel):
    # depthwise "same" filtering with zero padding: F.conv2d(..., padding=ws//2, groups=C), loss.py:100-108
    return F.conv2d(x, kernel, padding=kernel.shape[-1] // 2, groups=x.size(-3))


def ssim(img1, img2, window_size=11, size_average=True):
    """ssim() + _ssim() + create_window() of pointrix/model/loss.py:73-123 as one function: the 2-D window is
    the fp32 outer product of the 1-D window, one copy per channel; the SSIM map is
    (2 mu_x mu_y + C1)(2 cov + C2) / ((mu_x^2 + mu_y^2 + C1)(var_x + var_y + C2)) with C1 = 0.01^2, C2 = 0.03^2."""
    n_ch = img1.size(-3)
    w = window_1d(window_size)
    # window.type_as(img1) in the reference (loss.py:91-95): device AND dtype follow the image
    kernel = torch.outer(w, w).to(device=img1.device, dtype=img1.dtype).expand(n_ch, 1, window_size, window_size).contiguous()
    mean_x, mean_y = _box(img1, kernel), _box(img2, kernel)
    var_x = _box(img1 * img1, kernel) - mean_x * mean_x
    var_y = _box(img2 * img2, kernel) - mean_y * mean_y
    cov = _box(img1 * img2, kernel) - mean_x * mean_y
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    numer = (2 * (mean_x * mean_y) + c1) * (2 * cov + c2)
    denom = (mean_x * mean_x + mean_y * mean_y + c1) * (var_x + var_y + c2)
    smap = numer / denom
    if size_average:
        return smap.mean()
    return smap.mean(1).mean(1).mean(1)  # per image, the reference's nested means (loss.py:117)


def l1_ssim_loss(pred, gt, lambda_ssim=0.2):
    """BaseModel.get_loss_dict, pointrix/model/base_model.py:113-124."""
    l1 = l1_loss(pred, gt)
    one_minus_ssim = 1.0 - ssim(pred, gt)
    total = (1.0 - lambda_ssim) * l1 + lambda_ssim * one_minus_ssim
    return {"loss": total, "L1_loss": l1, "ssim_loss": one_minus_ssim}
