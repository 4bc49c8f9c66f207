#!/usr/bin/env python
"""Developer tool: pair-test statistics of the blend kernels on one cfg view.

Needs the instrumented library (`python pointrix_b200/csrc/build.py --stats`), loaded through
PXB_LIBRARY; prints how many (8x4 block, Gaussian) candidates survive the block mask, how many of
them blend at least one pixel, and the mean number of blending lanes.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PXB_LIBRARY"] = os.path.join(ROOT, "pointrix_b200", "libpointrix_b200_stats.so")

import torch  # noqa: E402

import pointrix_b200 as pb  # noqa: E402
from pointrix_b200 import _lib, scene  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
c, sc, cams = scene.make_config(cfg)
dev = torch.device("cuda", 0)
params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
cams = {k: v.to(dev) for k, v in cams.items()}
dimg = scene.upstream_gradient(3, c["H"], c["W"]).to(dev)
r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device="cuda:0")
r.sh_degree = 3
fn = _lib.lib.pxb_blend_stats
fn.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 8)()
res = []
for v in range(3):
    out = r.render_iter(c["H"], c["W"], cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    if v == 0:
        fn(buf, 1)  # first view also runs the exact binning path: discard
        out = r.render_iter(c["H"], c["W"], cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    fn(buf, 0)
    fwd_cand = buf[5]
    (out["rendered_features_split"]["rgb"] * dimg).sum().backward()
    fn(buf, 1)
    N = int((out["radii"] > 0).sum())
    res.append({"view": v, "visible": N, "fwd_block_candidates": fwd_cand, "bwd_block_candidates": buf[0],
                "bwd_any_valid": buf[1], "bwd_valid_lanes": buf[2], "bwd_flushes": buf[3], "bwd_partial_flush_slots": buf[4],
                "lanes_per_any_valid": buf[2] / max(1, buf[1])})
print(json.dumps(res, indent=1))
