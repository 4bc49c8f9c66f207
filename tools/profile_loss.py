#!/usr/bin/env python
"""Developer tool: a few fused L1+SSIM loss fwd+bwd at 1080p for `ncu --set full -k regex:l1_ssim`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pointrix_b200 import loss as PL  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(2)
gts = [torch.rand(1, 3, 1080, 1920, device="cuda", generator=g) for _ in range(4)]
preds = [(t + 0.1 * torch.randn(t.shape, device="cuda", generator=g)).clamp(0, 1) for t in gts]
for i in range(4):
    p = preds[i].detach().requires_grad_()
    PL.l1_ssim_loss(p, gts[i], 0.2)["loss"].backward()
torch.cuda.synchronize()
