#!/usr/bin/env python
"""Developer tool (torchrun): where the data-parallel step's exchange time goes, per rank:
fwd+bwd | radii copy-in | barrier 0 (= waiting for the slowest rank) | all-reduce kernel | barrier 1 | copy-out."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import pointrix_b200 as pb
from pointrix_b200 import _lib, parallel, renderer, scene

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
c, sc, cams = scene.make_config("cfg4")
H, W, P, V = c["H"], c["W"], c["P"], c["views"]
params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
cams = {k: v.to(dev) for k, v in cams.items()}
dimg = scene.upstream_gradient(3, H, W).to(dev) / world
r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device=str(dev))
r.sh_degree = 3
mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
factored = len(sys.argv) > 2 and sys.argv[2] == "factored"
ex = (parallel.ShFactoredExchange if factored else parallel.NvlsGradExchange)(P, dev, mode=mode)
renderer.set_grad_sink(ex)
import ctypes as C

ev = lambda: torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for v in range(V):
        r.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
names = ["fwd_bwd", "copy_in", "barrier0", "kernel", "gather", "barrier1", "copy_out"]
acc = {k: 0.0 for k in names}
K = 40
tot = 0.0
for it in range(5 + K):
    v = (it * world + rank) % V
    for p in params.values():
        p.grad = None
    e = [ev() for _ in range(8)]
    e[0].record()
    out = r.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params)
    (out["rendered_features_split"]["rgb"] * dimg).sum().backward()
    e[1].record()
    # NvlsGradExchange.exchange, phase by phase
    buf, h = ex.bufs[ex._cur], ex.hdls[ex._cur]
    cur, ex._cur = ex._cur, None
    radii = out["radii"]
    ri = buf[ex.n_f32:ex.n_f32 + ex.n_i32].view(torch.int32)
    pend, ex._pending = getattr(ex, "_pending", None), None
    ri[:P].copy_(radii.reshape(-1))
    e[2].record()
    h.barrier(channel=0)
    e[3].record()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    if ex.mode.startswith("nvls"):
        _lib.launch("pxb_nvls_allreduce", C.c_void_p(h.multicast_ptr), ex.n_f32, ex.n_i32, rank, world, st)
    else:
        _lib.launch("pxb_p2p_allreduce", ex._peer_arrays[cur], ex.n_f32, ex.n_i32, rank, world, st)
    e[4].record()
    if factored:
        _lib.launch("pxb_sh_grad_gather", ex._peer_arrays[cur], ex.rgb_off, ex.cam_off, world, P, pend[1],
                    C.c_void_p(params["position"].data_ptr()), C.c_void_p(pend[0].data_ptr()), st)
    e[5].record()
    h.barrier(channel=1)
    e[6].record()
    radii.reshape(-1).copy_(ri[:P])
    e[7].record()
    torch.cuda.synchronize()
    if it >= 5:
        for k, nm in enumerate(names):
            acc[nm] += e[k].elapsed_time(e[k + 1])
        tot += e[0].elapsed_time(e[7])
res = {k: round(v / K, 3) for k, v in acc.items()}
res["total"] = round(tot / K, 3)
gathered = [None] * world
dist.all_gather_object(gathered, res)
if rank == 0:
    print(f"mode={ex.mode} world={world}")
    for q, g in enumerate(gathered):
        print(q, g)
dist.destroy_process_group()
