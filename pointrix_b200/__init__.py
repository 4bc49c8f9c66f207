"""pointrix_b200 -- B200-native (sm_100a) msplat render path behind pointrix's
``MsplatRender`` plugin API.  See DESIGN.md / INTEGRATION.md.

Importing this package loads ``libpointrix_b200.so`` (hand-written CUDA behind
a C ABI, ``include/pointrix_b200.h``); there is no CPU or PyTorch fallback.
"""
from .ops import (  # noqa: F401
    alpha_blending,
    compute_cov3d,
    compute_sh,
    ewa_project,
    project_point,
    rasterization,
    sort_gaussian,
)
from . import densify  # noqa: F401  (clone / split / prune / opacity reset, pointrix/controller/gs.py)
from . import io  # noqa: F401  (.ply / .pth formats of the Gaussian table, pointrix/model/point_cloud/points.py:359-427)
from . import loss  # noqa: F401  (l1_loss / l2_loss / psnr / ssim / l1_ssim_loss, pointrix/model/loss.py)
from . import optim  # noqa: F401  (fused Adam + densification statistics, pointrix/optimizer/optimizer.py, controller/gs.py)
from .registry import RENDERER_REGISTRY, parse_renderer  # noqa: F401
from .renderer import MsplatRender, RenderFeatures, camera_extrinsics, fused_render  # noqa: F401

__all__ = [
    "project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "compute_sh", "alpha_blending",
    "rasterization", "MsplatRender", "RenderFeatures", "RENDERER_REGISTRY", "parse_renderer", "fused_render", "camera_extrinsics", "loss", "io", "optim", "densify",
]
