#!/usr/bin/env python
"""Developer tool (plain python or torchrun): host/device time of the bench step's parts."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import pointrix_b200 as pb
from pointrix_b200 import parallel, scene

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
c, sc, cams = scene.make_config("cfg4")
H, W = c["H"], c["W"]
params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
cams = {k: v.to(dev) for k, v in cams.items()}
dimg = scene.upstream_gradient(3, H, W).to(dev)
r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device=str(dev))
r.sh_degree = 3
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {"host_render": 0, "host_bwd": 0, "host_ar": 0, "dev_render": 0, "dev_bwd": 0, "dev_ar": 0, "dev_total": 0}
K = 20
for it in range(5 + K):
    v = (it * world + rank) % c["views"]
    for p in params.values():
        p.grad = None
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter(); e[0].record()
    out = r.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    t1 = time.perf_counter(); e[1].record()
    (out["rendered_features_split"]["rgb"] * dimg).sum().backward()
    t2 = time.perf_counter(); e[2].record()
    if world > 1:
        parallel.allreduce_step([p.grad for p in params.values()], out["uv_points"].grad, out["radii"], world)
    t3 = time.perf_counter(); e[3].record()
    torch.cuda.synchronize()
    if it >= 5:
        acc["host_render"] += (t1 - t0) * 1e3; acc["host_bwd"] += (t2 - t1) * 1e3; acc["host_ar"] += (t3 - t2) * 1e3
        acc["dev_render"] += e[0].elapsed_time(e[1]); acc["dev_bwd"] += e[1].elapsed_time(e[2]); acc["dev_ar"] += e[2].elapsed_time(e[3])
        acc["dev_total"] += e[0].elapsed_time(e[3])
if rank == 0:
    print({k: round(v / K, 3) for k, v in acc.items()}, file=sys.stderr)
if world > 1:
    dist.destroy_process_group()
