#!/usr/bin/env python
"""Developer tool: where the HOST time of one bench step goes (cProfile over un-synchronised steps),
next to the synchronised wall time per step."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import pointrix_b200 as pb
from pointrix_b200 import scene

dev = torch.device("cuda", 0)
c, sc, cams = scene.make_config("cfg4")
H, W = c["H"], c["W"]
params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
cams = {k: v.to(dev) for k, v in cams.items()}
dimg = scene.upstream_gradient(3, H, W).to(dev)
r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device=str(dev))
r.sh_degree = 3


def step(it):
    v = it % c["views"]
    for p in params.values():
        p.grad = None
    out = r.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    (out["rendered_features_split"]["rgb"] * dimg).sum().backward()


for it in range(8):
    step(it)
torch.cuda.synchronize()
K = 100
t0 = time.perf_counter()
for it in range(K):
    step(it)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"un-synchronised loop: host returns after {t_host / K * 1e3:.3f} ms/step, device done after {t_all / K * 1e3:.3f} ms/step")
# host-only cost: tiny scene (kernels are microseconds), same code path
c2, sc2, cams2 = scene.make_config("cfg1", P=2000, views=4)
p2 = {k: v.to(dev).requires_grad_() for k, v in sc2.items()}
cm2 = {k: v.to(dev) for k, v in cams2.items()}
d2 = scene.upstream_gradient(3, 64, 64).to(dev)


def step_small(it):
    v = it % 4
    for p in p2.values():
        p.grad = None
    out = r.render_iter(64, 64, cm2["extrinsic_matrix"][v], cm2["intrinsic_params"], cm2["camera_center"][v], **p2)
    (out["rendered_features_split"]["rgb"] * d2).sum().backward()


for it in range(8):
    step_small(it)
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(200):
    step_small(it)
torch.cuda.synchronize()
print(f"tiny scene (host-bound) : {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for it in range(200):
    step_small(it)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:6000])
print("cpu count", os.cpu_count())
