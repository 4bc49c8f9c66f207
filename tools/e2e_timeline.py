#!/usr/bin/env python
"""Developer tool: GPU timeline of ONE steady-state end-to-end step (bench.py's e2e_step): every kernel /
memcpy with its start offset, duration and the idle gap in front of it, from torch.profiler (CUPTI)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import pointrix_b200 as pb
from pointrix_b200 import scene

dev = torch.device("cuda", 0)
c, sc, cams = scene.make_config("cfg4")
H, W = c["H"], c["W"]
params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
V = c["views"]
host_cam = torch.cat([cams["extrinsic_matrix"].reshape(V, 16), cams["intrinsic_params"].reshape(1, 4).expand(V, 4),
                      cams["camera_center"].reshape(V, 3)], dim=1).contiguous().pin_memory()
host_dimg = scene.upstream_gradient(3, H, W).pin_memory()
res_host = torch.zeros(()).pin_memory()
r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device=str(dev))
r.sh_degree = 3
copy_stream = torch.cuda.Stream(device=dev)


def e2e_step(step):
    v = step % c["views"]
    main = torch.cuda.current_stream(dev)
    cam = host_cam[v].to(dev, non_blocking=True)
    E, I, Cc = cam[0:16].view(4, 4), cam[16:20], cam[20:23]
    copy_stream.wait_stream(main)
    with torch.cuda.stream(copy_stream):
        G = host_dimg.to(dev, non_blocking=True)
    for p_ in params.values():
        p_.grad = None
    out = r.render_iter(H, W, E, I, Cc, **params)
    img = out["rendered_features_split"]["rgb"]
    main.wait_stream(copy_stream)
    G.record_stream(main)
    loss = (img * G).sum()
    loss.backward()
    res_host.copy_(loss.detach(), non_blocking=True)
    main.synchronize()
    return float(res_host)


for s in range(6):
    e2e_step(s)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(4):
        e2e_step(6 + s)
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "e2e_trace.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
tr = json.load(open(path))
ev = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# split into steps at the D2H copy that ends each one
steps, cur = [], []
for e in ev:
    cur.append(e)
    if e["cat"] == "gpu_memcpy" and "DtoH" in e["name"] and e["args"].get("bytes", 0) == 4:
        steps.append(cur)
        cur = []
st = steps[2]
t0 = st[0]["ts"]
end = 0
busy = 0
print(f"{'start_us':>9} {'dur_us':>8} {'gap_us':>7}  name")
for e in st:
    gap = e["ts"] - end if end else 0
    print(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} {gap:7.1f}  {e['name'][:70]} [s{e['args'].get('stream')}]")
    end = max(end, e["ts"] + e["dur"])
    busy += e["dur"]
print(f"step GPU span {end - t0:.1f} us, sum of durations {busy:.1f} us")
# host span of the same step: from the first cudaMemcpyAsync to the stream synchronize
cpu = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver")]
cpu.sort(key=lambda e: e["ts"])
sy = [e for e in cpu if "StreamSynchronize" in e["name"]]
if len(sy) >= 3:
    print(f"host: step period {sy[2]['ts'] + sy[2]['dur'] - (sy[1]['ts'] + sy[1]['dur']):.1f} us; "
          f"final sync waited {sy[2]['dur']:.1f} us; first GPU op at +{t0 - (sy[1]['ts'] + sy[1]['dur']):.1f} us after the previous sync returned")
os.remove(path)
