#!/usr/bin/env python
"""Developer tool: run warm-up views, then ONE profiled render fwd+bwd (cfg4 by default) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full ...`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pointrix_b200 as pb  # noqa: E402
from pointrix_b200 import scene  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
c, sc, cams = scene.make_config(cfg)
dev = torch.device("cuda", 0)
params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
cams = {k: v.to(dev) for k, v in cams.items()}
dimg = scene.upstream_gradient(3, c["H"], c["W"]).to(dev)
r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device="cuda:0")
r.sh_degree = 3


def step(v):
    for p in params.values():
        p.grad = None
    out = r.render_iter(c["H"], c["W"], cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    (out["rendered_features_split"]["rgb"] * dimg).sum().backward()


for v in range(3):
    step(v)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
