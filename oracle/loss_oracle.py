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
    # gaussian(), pointrix/model/loss.py:69-71: python-double exp, fp32 tensor, fp32 normalisation
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


def ssim(img1, img2, window_size=11, size_average=True):
    # ssim() + _ssim() + create_window(), pointrix/model/loss.py:73-123
    channel = img1.size(-3)
    w1 = window_1d(window_size).unsqueeze(1)
    window = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()
    window = window.type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    if size_average:
        return ssim_map.mean()
    return ssim_map.mean(1).mean(1).mean(1)


def l1_ssim_loss(pred, gt, lambda_ssim=0.2):
    # BaseModel.get_loss_dict, pointrix/model/base_model.py:113-124
    L1 = l1_loss(pred, gt)
    sl = 1.0 - ssim(pred, gt)
    return {"loss": (1.0 - lambda_ssim) * L1 + lambda_ssim * sl, "L1_loss": L1, "ssim_loss": sl}
