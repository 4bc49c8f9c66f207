#!/usr/bin/env python
"""Every BASELINE.json config on ONE GPU, ours beside the compiled reference CUDA (oracle/_ref), CUDA events,
>= 5 warm-up steps, a different view every step.  One JSON object per config on stdout.

  cfg1  10K Gaussians, 800x800, rgb, fwd+bwd
  cfg2  300K Gaussians, 800x800, training step = render + 0.8 L1 + 0.2 (1-SSIM) loss + backward
  cfg3  3M Gaussians, 1297x840, fwd+bwd (per GPU; the 8-GPU run shards views)
  cfg4  1M Gaussians, 1920x1080, forward only, rgb+depth+normal+flow (C = 3+1+3+2 = 9) -> Mpix/s
  cfg5  1M Gaussians, 979x546, fwd+bwd with camera gradients (intrinsics, extrinsics, centre)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pointrix_b200 as pb  # noqa: E402
from oracle import ref_driver  # noqa: E402  (the compiled reference beside ours; never on the product path)
from pointrix_b200 import loss as PL  # noqa: E402
from pointrix_b200 import scene  # noqa: E402

dev = torch.device("cuda", 0)
K = int(os.environ.get("SWEEP_STEPS", 20))


def timed(fn, n=K, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, mode):
    c, sc, cams = scene.make_config(name, views=min(scene.CONFIGS[name]["views"], 16))
    H, W, P, V = c["H"], c["W"], c["P"], c["views"]
    params = {k: v.to(dev).requires_grad_(mode != "fwd9") for k, v in sc.items()}
    cams = {k: v.to(dev) for k, v in cams.items()}
    depth_ch = mode == "fwd9"
    r = pb.parse_renderer({"name": "MsplatRender", "render_depth": depth_ch}, white_bg=True, device="cuda:0")
    r.sh_degree = 3
    g = torch.Generator().manual_seed(7)
    extra = {}
    if mode == "fwd9":
        extra = {"normals": torch.randn(P, 3, generator=g).to(dev), "extra_features": {"flow": torch.randn(P, 2, generator=g).to(dev)}}
    C = 9 if mode == "fwd9" else 3
    dimg = scene.upstream_gradient(C, H, W).to(dev)
    gt = torch.rand(1, 3, H, W, generator=g).to(dev)
    with torch.no_grad():
        for v in range(V):
            r.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **params, **extra)

    def ours(i, bwd=True):
        v = i % V
        E, I, Cc = cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v]
        if mode == "camgrad":
            E, I, Cc = (t.clone().requires_grad_() for t in (E, I, Cc))
        for p_ in params.values():
            p_.grad = None
        if not bwd or mode == "fwd9":
            with torch.no_grad():
                r.render_iter(H, W, E, I, Cc, **params, **extra)
            return
        out = r.render_iter(H, W, E, I, Cc, **params)
        img = out["rendered_features_split"]["rgb"]
        if mode == "loss":
            PL.l1_ssim_loss(img.unsqueeze(0), gt, 0.2)["loss"].backward()
        else:
            (img * dimg).sum().backward()

    def ref(i, bwd=True):
        v = i % V
        scd = {k: t.detach() for k, t in params.items()}
        ex = torch.cat(list(extra.values()), dim=1) if extra else None
        f = ref_driver.render_forward(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v],
                                      **scd, render_depth=depth_ch, extra=ex)
        if not bwd or mode == "fwd9":
            return
        if mode == "loss":
            from oracle import loss_oracle as LO

            im = f["img"].detach().requires_grad_()
            LO.l1_ssim_loss(im.unsqueeze(0), gt, 0.2)["loss"].backward()
            d = im.grad
        else:
            d = dimg
        ref_driver.render_backward(f, d, scd["position"], scd["opacity"], scd["scaling"], scd["rotation"], scd["shs"],
                                   cams["camera_center"][v], camera_grads=(mode == "camgrad"))

    res = {"config": name, "P": P, "W": W, "H": H, "channels": C, "mode": mode}
    ms_f = timed(lambda i: ours(i, False))
    res["ours_fwd_ms"] = round(ms_f, 4)
    res["ours_render_mpix_s"] = round(W * H / 1e6 / (ms_f / 1e3), 1)
    if mode != "fwd9":
        ms = timed(ours)
        res["ours_fwd_bwd_ms"], res["ours_it_s"] = round(ms, 4), round(1e3 / ms, 1)
    if ref_driver.available():
        n_ref = max(3, K // 4)
        rf = timed(lambda i: ref(i, False), n=n_ref, warm=2)
        res["ref_fwd_ms"], res["ref_render_mpix_s"] = round(rf, 4), round(W * H / 1e6 / (rf / 1e3), 1)
        if mode != "fwd9":
            rb = timed(ref, n=n_ref, warm=2)
            res["ref_fwd_bwd_ms"], res["ref_it_s"] = round(rb, 4), round(1e3 / rb, 1)
    from pointrix_b200 import ops

    res["intersections_binned_last_view"] = int(ops.LAST_N.get((0, True), -1))
    print(json.dumps(res), flush=True)
    del params, cams
    torch.cuda.empty_cache()


for name, mode in (("cfg1", "rgb"), ("cfg2", "loss"), ("cfg3", "rgb"), ("cfg4", "fwd9"), ("cfg5", "camgrad")):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    run(name, mode)
