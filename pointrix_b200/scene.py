"""Deterministic synthetic scenes and cameras (BASELINE.md section 3 / SURVEY.md 8d).

Everything is drawn from a CPU ``torch.Generator`` and then copied to the
device, so CPU (oracle) and GPU runs see bit-identical inputs.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

# (P, W, H, views, s_lo, s_hi) per BASELINE.json config
CONFIGS = {
    "cfg1": dict(P=10_000, W=800, H=800, views=1, s_lo=0.01, s_hi=0.08),
    "cfg2": dict(P=300_000, W=800, H=800, views=100, s_lo=0.004, s_hi=0.03),
    "cfg3": dict(P=3_000_000, W=1297, H=840, views=200, s_lo=0.002, s_hi=0.015),
    "cfg4": dict(P=1_000_000, W=1920, H=1080, views=16, s_lo=0.003, s_hi=0.02),
    "cfg5": dict(P=1_000_000, W=979, H=546, views=64, s_lo=0.003, s_hi=0.02),
}


def make_scene(P: int, s_lo: float, s_hi: float, seed: int = 0, sh_k: int = 16, device="cpu") -> Dict[str, torch.Tensor]:
    """Post-activation Gaussian table with pointrix's shapes
    (pointrix/model/point_cloud/gaussian_points.py:70-86): position[P,3],
    scaling[P,3], rotation[P,4] (unit, w first), opacity[P,1], shs[P,sh_k,3]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    pos = (torch.rand(P, 3, generator=g) * 2.0 - 1.0) * 1.3
    log_s = torch.rand(P, 3, generator=g) * (math.log(s_hi) - math.log(s_lo)) + math.log(s_lo)
    quat = torch.randn(P, 4, generator=g)
    quat = quat / quat.norm(dim=1, keepdim=True)
    opacity = torch.sigmoid(torch.randn(P, 1, generator=g) * 2.0)
    shs = torch.randn(P, sh_k, 3, generator=g) * 0.1
    shs[:, 0, :] = torch.rand(P, 3, generator=g) * 3.0 - 1.5
    out = dict(position=pos, scaling=torch.exp(log_s), rotation=quat, opacity=opacity, shs=shs)
    return {k: v.float().contiguous().to(device) for k, v in out.items()}


def look_at_camera(eye: torch.Tensor) -> torch.Tensor:
    """4x4 world->camera matrix, OpenCV convention (x right, y down, z forward), looking at the origin."""
    eye = eye.double()
    fwd = -eye / eye.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    if abs(float(fwd @ up)) > 0.999:
        up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    R = torch.stack([right, down, fwd], dim=0)
    E = torch.eye(4, dtype=torch.float64)
    E[:3, :3] = R
    E[:3, 3] = -R @ eye
    return E.float()


def make_cameras(n: int, W: int, H: int, seed: int = 1, radius: float = 4.03, device="cpu") -> Dict[str, torch.Tensor]:
    """Blender-style orbit: extrinsic_matrix[n,4,4], camera_center[n,3], intrinsic_params[4]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    az = torch.rand(n, generator=g) * 2.0 * math.pi
    el = (torch.rand(n, generator=g) * 0.8 + 0.1) * (math.pi / 2.0) * 0.9
    eyes = torch.stack([torch.cos(az) * torch.cos(el), torch.sin(az) * torch.cos(el), torch.sin(el)], -1) * radius
    E = torch.stack([look_at_camera(e) for e in eyes], 0)
    f = 0.5 * W / math.tan(0.5 * 0.6911)
    intr = torch.tensor([f, f, W / 2.0, H / 2.0], dtype=torch.float32)
    return dict(extrinsic_matrix=E.to(device), camera_center=eyes.float().to(device), intrinsic_params=intr.to(device))


def make_config(name: str, device="cpu", P: int | None = None, views: int | None = None) -> Tuple[Dict, Dict, Dict]:
    c = dict(CONFIGS[name])
    if P is not None:
        c["P"] = P
    if views is not None:
        c["views"] = views
    scene = make_scene(c["P"], c["s_lo"], c["s_hi"], seed=0, device=device)
    cams = make_cameras(c["views"], c["W"], c["H"], seed=1, device=device)
    return c, scene, cams


def upstream_gradient(C: int, H: int, W: int, seed: int = 2, device="cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(C, H, W, generator=g).float().to(device)


def target_images(n: int, H: int, W: int, seed: int = 2) -> torch.Tensor:
    """``n`` fixed random target images [n,3,H,W] in [0,1] (the ground truth of the photometric loss,
    SURVEY.md 8d: seed 2 = upstream-gradient / target images)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.rand(n, 3, H, W, generator=g).float()


def extra_features(P: int, seed: int = 3) -> Dict[str, torch.Tensor]:
    """Per-Gaussian feature columns of the cfg4 "render anything" sweep beside rgb and depth: unit
    ``normals[P,3]`` and a 2-channel ``flow[P,2]`` (optical flow is representable only as
    generic feature channels), passed to ``render_iter`` as keyword arguments => C = 3 + 1 + 3 + 2 = 9."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = torch.randn(P, 3, generator=g)
    n = n / n.norm(dim=1, keepdim=True)
    return {"normals": n.float().contiguous(), "flow": (torch.randn(P, 2, generator=g) * 2.0).float().contiguous()}
