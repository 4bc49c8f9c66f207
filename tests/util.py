"""Shared helpers for the parity tests."""
import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max|b|  -- the 'relative 1e-3 on gradients' measure of the north star."""
    a, b = a.double().cpu(), b.double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / max(den, 1e-30)


def l2_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def to_dev(d, device):
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def scene_inputs(name="cfg1", P=None, views=1, device="cuda", W=None, H=None):
    from pointrix_b200 import scene

    c, sc, cams = scene.make_config(name, P=P, views=views)
    if W is not None:
        c["W"], c["H"] = W, H
        cams = scene.make_cameras(views, W, H, seed=1)
    return c, to_dev(sc, device), to_dev(cams, device)


def elem_bad_fraction(a: torch.Tensor, b: torch.Tensor, rtol: float = 1e-3) -> float:
    """Per-element reading of 'relative 1e-3 on gradients': the fraction of elements with
    |a - b| > rtol * |b| + rtol * typical(|b|), typical = the median magnitude of the non-zero reference
    entries (the absolute floor every fp32 sum of signed terms needs; the max-normalised rel_err above
    uses the LARGEST entry as that floor and so barely looks at small gradients)."""
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    nz = b.abs()[b != 0]
    floor = nz.median().item() if nz.numel() else 0.0
    bad = (a - b).abs() > rtol * b.abs() + rtol * floor
    return bad.double().mean().item()
