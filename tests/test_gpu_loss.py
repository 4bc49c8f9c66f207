"""GPU parity of the fused photometric loss (csrc/loss.cu) through the public
Python API (= through the C ABI) against
  (a) the committed golden vectors produced by the reference's own pointrix/model/loss.py
      (tests/golden/ref_loss.npz), and
  (b) the CPU oracle (oracle/loss_oracle.py) on seeded inputs incl. ragged / tiny / full-size images.
Tolerances (floating point; the separable filter differs from the reference's 2-D cuDNN/CPU convolution
only in fp32 summation order): loss values max-abs 1e-5, per-pixel maps 1e-6, gradients relative 1e-3
(north star) -- measured errors are ~1e-5 relative.
"""
import os

import numpy as np
import pytest
import torch

from tests.util import l2_rel, rel_err

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-5
GRAD_TOL = 1e-3


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from pointrix_b200 import loss

    return loss


@pytest.fixture(scope="module")
def GL():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_loss.npz"))


def _c(a):
    return torch.tensor(a).cuda()


@pytest.mark.parametrize("tag", list("abcde"))
def test_loss_vs_reference_golden(L, GL, tag):
    pred, gt = _c(GL[f"{tag}_pred"]), _c(GL[f"{tag}_gt"])
    # l1 / l2 (+ unreduced map) and their gradients
    p = pred.clone().requires_grad_()
    v = L.l1_loss(p, gt)
    v.backward()
    assert abs(v.item() - float(GL[f"{tag}_l1"])) <= VAL_TOL
    assert rel_err(p.grad, torch.tensor(GL[f"{tag}_l1_grad"])) <= 1e-6
    p = pred.clone().requires_grad_()
    v = L.l2_loss(p, gt)
    v.backward()
    assert abs(v.item() - float(GL[f"{tag}_l2"])) <= VAL_TOL
    assert rel_err(p.grad, torch.tensor(GL[f"{tag}_l2_grad"])) <= 1e-5
    assert torch.equal(L.l1_loss(pred, gt, return_mean=False).cpu(), torch.tensor(GL[f"{tag}_l1_map"]))
    if pred.dim() == 4:
        assert torch.allclose(L.psnr(pred, gt).cpu(), torch.tensor(GL[f"{tag}_psnr"]), atol=1e-4)
        assert torch.allclose(L.ssim(pred, gt, size_average=False).cpu(), torch.tensor(GL[f"{tag}_ssim_per_image"]), atol=VAL_TOL)
    # ssim value + gradient
    p = pred.clone().requires_grad_()
    s = L.ssim(p, gt)
    s.backward()
    assert abs(s.item() - float(GL[f"{tag}_ssim"])) <= VAL_TOL
    assert rel_err(p.grad, torch.tensor(GL[f"{tag}_ssim_grad"])) <= GRAD_TOL
    # the get_loss_dict combination from one pass
    p = pred.clone().requires_grad_()
    d = L.l1_ssim_loss(p, gt, 0.2)
    d["loss"].backward()
    assert abs(d["loss"].item() - float(GL[f"{tag}_loss"])) <= VAL_TOL
    assert abs(d["L1_loss"].item() - float(GL[f"{tag}_l1"])) <= VAL_TOL
    assert abs(d["ssim_loss"].item() - (1.0 - float(GL[f"{tag}_ssim"]))) <= VAL_TOL
    assert rel_err(p.grad, torch.tensor(GL[f"{tag}_loss_grad"])) <= GRAD_TOL


@pytest.mark.parametrize("shape", [(1, 3, 800, 800), (2, 3, 131, 257), (1, 1, 5, 300), (1, 3, 32, 32), (1, 3, 33, 31), (4, 2, 11, 11)])
def test_loss_vs_cpu_oracle(L, shape):
    from oracle import loss_oracle as LO

    g = torch.Generator().manual_seed(shape[-1])
    gt = torch.rand(shape, generator=g)
    pred = (gt + 0.2 * torch.randn(shape, generator=g)).clamp(0, 1)
    po = pred.clone().requires_grad_()
    do = LO.l1_ssim_loss(po, gt, 0.2)
    do["loss"].backward()
    pc = pred.cuda().requires_grad_()
    dc = L.l1_ssim_loss(pc, gt.cuda(), 0.2)
    dc["loss"].backward()
    for k in ("loss", "L1_loss", "ssim_loss"):
        assert abs(dc[k].item() - do[k].item()) <= VAL_TOL, k
    assert rel_err(pc.grad, po.grad) <= GRAD_TOL and l2_rel(pc.grad, po.grad) <= 1e-4
    # per-image SSIM with a non-uniform upstream gradient (exercises the [B] weight path)
    wts = torch.arange(1, shape[0] + 1, dtype=torch.float32)
    po = pred.clone().requires_grad_()
    (LO.ssim(po, gt, size_average=False) * wts).sum().backward()
    pc = pred.cuda().requires_grad_()
    (L.ssim(pc, gt.cuda(), size_average=False) * wts.cuda()).sum().backward()
    assert rel_err(pc.grad, po.grad) <= GRAD_TOL


def test_loss_dict_other_heads_and_upstream_scale(L):
    """Backward through L1_loss / ssim_loss heads and a non-unit upstream gradient of `loss`."""
    from oracle import loss_oracle as LO

    g = torch.Generator().manual_seed(3)
    gt, pred = torch.rand(2, 3, 45, 70, generator=g), torch.rand(2, 3, 45, 70, generator=g)
    po = pred.clone().requires_grad_()
    do = LO.l1_ssim_loss(po, gt, 0.35)
    (3.0 * do["loss"] + 0.5 * do["L1_loss"] - 2.0 * do["ssim_loss"]).backward()
    pc = pred.cuda().requires_grad_()
    dc = L.l1_ssim_loss(pc, gt.cuda(), 0.35)
    (3.0 * dc["loss"] + 0.5 * dc["L1_loss"] - 2.0 * dc["ssim_loss"]).backward()
    assert rel_err(pc.grad, po.grad) <= GRAD_TOL
    pc2 = pred.cuda().requires_grad_()
    (7.0 * L.l1_ssim_loss(pc2, gt.cuda(), 0.35)["loss"]).backward()
    po2 = pred.clone().requires_grad_()
    (7.0 * LO.l1_ssim_loss(po2, gt, 0.35)["loss"]).backward()
    assert rel_err(pc2.grad, po2.grad) <= GRAD_TOL


def test_loss_full_size_properties(L):
    """1080p (BASELINE cfg4 image size): size-independent properties instead of a CPU run."""
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(1, 3, 1080, 1920, device="cuda", generator=g)
    y = torch.rand(1, 3, 1080, 1920, device="cuda", generator=g)
    assert abs(L.ssim(x, x).item() - 1.0) <= 1e-6                      # identity
    assert L.l1_loss(x, x).item() == 0.0
    assert abs(L.ssim(x, y).item() - L.ssim(y, x).item()) <= 1e-7       # symmetry
    a, b = L.l1_ssim_loss(x, y, 0.2), L.l1_ssim_loss(x, y, 0.2)
    assert a["loss"].item() == b["loss"].item()                         # deterministic reduction
    # batch of two images == the two images separately (means of equal-sized images)
    xb, yb = torch.cat([x, y]), torch.cat([y, x])
    s2 = L.ssim(xb, yb, size_average=False)
    assert abs(s2[0].item() - L.ssim(x, y).item()) <= 1e-7 and abs(s2[1].item() - s2[0].item()) <= 1e-7
    # gradient of the loss is a descent direction: a small step along -grad lowers it
    p = x.clone().requires_grad_()
    l0 = L.l1_ssim_loss(p, y, 0.2)["loss"]
    l0.backward()
    step = 0.5 * p.grad / p.grad.abs().max()
    l1 = L.l1_ssim_loss((x - 0.01 * step), y, 0.2)["loss"]
    pred_drop = (p.grad * 0.01 * step).sum().item()
    assert l1.item() < l0.item() and abs((l0.item() - l1.item()) - pred_drop) <= 0.2 * pred_drop
    # l1 map mean == l1 mean
    assert abs(L.l1_loss(x, y, return_mean=False).mean().item() - L.l1_loss(x, y).item()) <= 1e-6


def test_loss_errors(L):
    x = torch.rand(1, 3, 16, 16, device="cuda")
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        L.ssim(torch.rand(1, 3, 16, 16), torch.rand(1, 3, 16, 16))
    with pytest.raises(ValueError, match="window_size"):
        L.ssim(x, x, window_size=7)
    with pytest.raises(AssertionError):
        L.l1_loss(x, x[:, :2])
    with pytest.raises(NotImplementedError):
        L.ssim(x, x.clone().requires_grad_())
    # no autograd requested -> no derivative maps are written, values unchanged
    with torch.no_grad():
        a = L.l1_ssim_loss(x, x * 0.5, 0.2)["loss"].item()
    b = L.l1_ssim_loss(x.clone().requires_grad_(), x * 0.5, 0.2)["loss"].item()
    assert abs(a - b) <= 1e-6


def test_render_then_loss_end_to_end(L):
    """render_iter -> fused loss -> backward: gradients reach the Gaussians, and match the same graph with
    the loss evaluated by the CPU oracle on the rendered image (chain rule through dL/dimage)."""
    import pointrix_b200 as pb
    from oracle import loss_oracle as LO
    from tests.util import scene_inputs

    c, sc, cams = scene_inputs("cfg1", P=4000, W=200, H=152)
    r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device="cuda:0")
    r.sh_degree = 3
    gt = torch.rand(3, 152, 200, generator=torch.Generator().manual_seed(2))
    grads = []
    for which in ("cuda", "oracle"):
        leaves = {k: v.clone().requires_grad_() for k, v in sc.items()}
        out = r.render_iter(152, 200, cams["extrinsic_matrix"][0], cams["intrinsic_params"], cams["camera_center"][0], **leaves)
        img = out["rendered_features_split"]["rgb"]
        if which == "cuda":
            L.l1_ssim_loss(img.unsqueeze(0), gt.cuda().unsqueeze(0), 0.2)["loss"].backward()
        else:
            ic = img.detach().cpu().requires_grad_()
            LO.l1_ssim_loss(ic.unsqueeze(0), gt.unsqueeze(0), 0.2)["loss"].backward()
            img.backward(ic.grad.cuda())
        grads.append({k: v.grad.clone() for k, v in leaves.items()})
    for k in grads[0]:
        assert grads[0][k].abs().sum().item() > 0, k
        assert l2_rel(grads[0][k], grads[1][k]) <= GRAD_TOL, k


def test_loss_oracle_runs_on_cuda_tensors(L):
    """bench.py times oracle/loss_oracle.py (the reference's torch-op formulation) on CUDA tensors beside the
    fused kernels: its SSIM window must follow the image's device (`window.type_as(img1)`, loss.py:91-95)."""
    from oracle import loss_oracle as LO

    g = torch.Generator().manual_seed(5)
    gt = torch.rand(1, 3, 70, 90, generator=g).cuda()
    pred = (gt + 0.1 * torch.randn(gt.shape, generator=g).cuda()).clamp(0, 1)
    p1 = pred.clone().requires_grad_()
    d1 = LO.l1_ssim_loss(p1, gt, 0.2)
    d1["loss"].backward()
    p2 = pred.clone().requires_grad_()
    d2 = L.l1_ssim_loss(p2, gt, 0.2)
    d2["loss"].backward()
    assert abs(d1["loss"].item() - d2["loss"].item()) <= VAL_TOL
    assert rel_err(p2.grad, p1.grad) <= GRAD_TOL
