#!/usr/bin/env python
"""Developer tool: device-resident step time with / without the per-stage event timer, host time per step,
and the time the host spends spinning on the intersection count."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import pointrix_b200 as pb
from pointrix_b200 import _lib, renderer, scene

dev = torch.device("cuda", 0)
c, sc, cams = scene.make_config("cfg4")
H, W, V = c["H"], c["W"], c["views"]
params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
cams = {k: v.to(dev) for k, v in cams.items()}
dimg = scene.upstream_gradient(3, H, W).to(dev)
r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device=str(dev))
r.sh_degree = 3

spin = [0.0]
_orig_wait = renderer._wait_count


def _timed_wait(word):
    t0 = time.perf_counter()
    n = _orig_wait(word)
    spin[0] += time.perf_counter() - t0
    return n


def step(it):
    v = it % V
    for p in params.values():
        p.grad = None
    out = r.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    (out["rendered_features_split"]["rgb"] * dimg).sum().backward()


def run(K, label):
    for it in range(5):
        step(it)
    torch.cuda.synchronize()
    spin[0] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for it in range(K):
        step(5 + it)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{label}: device {e0.elapsed_time(e1) / K:.4f} ms/step, host loop {1e3 * t_host / K:.4f} ms/step, "
          f"of which spinning on the count {1e3 * spin[0] / K:.4f} ms/step", flush=True)


renderer._wait_count = _timed_wait
run(100, "no timer")
_lib.set_timer(_lib.KernelTimer())
run(100, "stage-event timer")
_lib.set_timer(None)
run(100, "no timer again")
import bench  # noqa: E402  (the clock sampler of the bench: does sampling perturb the step?)

smp = bench.ClockSampler(0)
smp.start()
smp.wait_first(3.0)
smp.mark()
run(100, "no timer, NVML clock sampler running")
smp.mark_end()
print("   sampler:", smp.stop(), flush=True)
_lib.set_timer(_lib.KernelTimer(stages={"pxb_blend_backward"}))
run(100, "dominant-kernel events only")
_lib.set_timer(None)
# forward only
with torch.no_grad():
    def fstep(it):
        v = it % V
        r.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    for it in range(5):
        fstep(it)
    torch.cuda.synchronize()
    spin[0] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for it in range(100):
        fstep(it)
    e1.record()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"forward only: device {e0.elapsed_time(e1) / 100:.4f} ms, host {10 * th:.4f} ms, spin {10 * spin[0]:.4f} ms")
