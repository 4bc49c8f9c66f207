"""Golden-vector generator -- TEST INFRASTRUCTURE ONLY.

Two sources, both the reference itself:

  python oracle/make_golden.py --from-ref-tests
      (this container, CPU) imports the reference's OWN in-test PyTorch oracles from
      /root/reference/msplat/test/test_*.py (project_point_torch_impl,
      compute_cov3d_torch_impl, ewa_project_torch_impl, eval_sh_bases,
      alpha_blending_torch_impl) with the reference's seeds / shapes, runs them on CPU
      and stores inputs + outputs (+ autograd gradients) in
      tests/golden/ref_test_oracles.npz.  Nothing is copied: the functions are imported
      where they lie, `msplat` is stubbed and `.cuda()` is made a no-op.

  python oracle/make_golden.py --from-ref-gpu [out.npz]
      (GPU box) runs the compiled unmodified reference CUDA (oracle/_ref) on a small
      seeded scene through its own op order and stores every intermediate and gradient
      -> tests/golden/ref_gpu_small.npz (written to gpurun_out/ on the box and copied).

  python oracle/make_golden.py --from-ref-loss
      (this container, CPU) imports the reference's OWN pointrix/model/loss.py where it lies (its
      `lpips` import, an absent third-party package the L1/SSIM code never touches, is stubbed),
      runs l1_loss / l2_loss / psnr / ssim and the get_loss_dict combination with autograd on small
      seeded images and stores inputs, values and gradients in tests/golden/ref_loss.npz.

  python oracle/make_golden.py --from-ref-plugin
      (this container, CPU) executes the reference's OWN plugin file pointrix/model/renderer/msplat.py where
      it lies, against a stub `BaseObject`, the reference's real registry.py / renderer_utils.py and a module
      `msplat` backed by the CPU oracle's six operators: render_iter (rgb / rgb+depth, SH degree 3 / 1,
      autograd gradients), render_batch and update_sh_degree -> tests/golden/ref_plugin.npz.  Pins the
      plugin glue (SH degree mask, +0.5 clamp, nearest = 0.2, depth channel, ndc.grad, batch reductions).

  python oracle/make_golden.py --from-ref-controller
      (this container, CPU) executes the reference's OWN pointrix/controller/gs.py where it lies (its imports
      of the config / pose / base-object modules stubbed) and calls DensificationController.preprocess --
      accumulate_viewspace_grad + the masked updates of grad_accum / acc_steps / max_radii -- three iterations
      on seeded inputs (two views per iteration) -> tests/golden/ref_controller.npz.

  python oracle/make_golden.py --from-ref-densify
      (this container, CPU) the clone / split / prune / opacity-reset surgery: the reference's OWN GaussianPointCloud
      (points.py, gaussian_points.py, point_utils.py), pose.py and DensificationController.densify (gs.py, base.py)
      executed where they lie on an 800-point cloud with a populated torch.optim.Adam, at steps 600 (clone + split +
      prune), 3000 (... + opacity reset) and 3100 (prune with the size thresholds) -> tests/golden/ref_densify.npz:
      the table, both Adam moments and the controller state before and after each call.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF_TESTS = "/root/reference/msplat/test"


def _import_ref_test(name):
    sys.modules.setdefault("msplat", types.ModuleType("msplat"))
    import scipy.special as sp

    if not hasattr(sp, "sph_harm"):  # removed in recent scipy; the test only needs the name at import
        sp.sph_harm = lambda m, n, theta, phi: sp.sph_harm_y(n, m, phi, theta)
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF_TESTS, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def from_ref_tests(out_path):
    torch.Tensor.cuda = lambda self, *a, **k: self  # the helpers hard-code .cuda()
    G = {}
    # ---- project_point (test_project_points.py:55-77 inputs)
    m = _import_ref_test("test_project_points")
    torch.manual_seed(123)
    W, H, N = 1600, 1200, 1000
    intr = torch.tensor([2892.33, 2883.18, 823.205, 619.071])
    extr = torch.tensor([[0.970263, 0.00747983, 0.241939, -191.02], [-0.0147429, 0.999493, 0.0282234, 3.2883],
                         [-0.241605, -0.030951, 0.969881, 22.5401]])
    xyz = torch.rand(N, 3)
    xyz[:, 0] *= 500; xyz[:, 1] *= 500; xyz[:, 2] = xyz[:, 2] * 400 + 400
    x1 = xyz.clone().requires_grad_()
    uv, depth = m.project_point_torch_impl(x1, intr, extr, W, H)
    (uv.sum() + depth.sum()).backward()
    G.update(proj_xyz=xyz, proj_intr=intr, proj_extr=extr, proj_WH=torch.tensor([W, H]), proj_uv=uv.detach(),
             proj_depth=depth.detach(), proj_dxyz=x1.grad)
    # ---- compute_cov3d (test_compute_cov3d.py:38-47)
    m = _import_ref_test("test_compute_cov3d")
    torch.manual_seed(7)
    N = 2000
    s = torch.rand(N, 3)
    q = torch.randn(N, 4); q = q / q.norm(dim=-1, keepdim=True)
    s1, q1 = s.clone().requires_grad_(), q.clone().requires_grad_()
    cov = m.compute_cov3d_torch_impl(s1, q1)
    cov.sum().backward()
    G.update(cov_scales=s, cov_quats=q, cov_out=cov.detach(), cov_ds=s1.grad, cov_dq=q1.grad)
    # ---- ewa_project (test_ewa_project.py:140-161)
    m = _import_ref_test("test_ewa_project")
    torch.manual_seed(123)
    N, W, H = 4000, 800, 800
    intr = torch.tensor([1111.0, 1111.0, H / 2, W / 2 / 2])
    extr = torch.tensor([[6.1182e-01, 7.9099e-01, 1.3906e-14, 1.1327e-09], [7.9096e-01, -6.1180e-01, -8.5126e-03, 1.0458e-09],
                         [-6.7348e-03, 5.2093e-03, -9.9996e-01, 4.0311e+00]])
    xyz = torch.randn(N, 3) * 2.6 - 1.3
    sc = torch.rand(N, 3) + 1
    qq = torch.rand(N, 4); qq = qq / qq.norm(dim=-1, keepdim=True)
    from oracle import msplat_oracle as O

    cov3d = O.compute_cov3d(sc, qq)
    uv, depth = O.project_point(xyz, intr, extr, W, H, nearest=0.2)
    vis = (depth != 0).reshape(-1)
    m.uv = uv  # the helper reads a module-global `uv` for its device
    x1, c1, i1, e1 = xyz.clone().requires_grad_(), cov3d.clone().requires_grad_(), intr.clone().requires_grad_(), extr.clone().requires_grad_()
    conic, radius, tiles = m.ewa_project_torch_impl(x1, c1, i1, e1, uv, W, H, vis)
    conic.sum().backward()
    G.update(ewa_xyz=xyz, ewa_cov3d=cov3d, ewa_intr=intr, ewa_extr=extr, ewa_uv=uv, ewa_vis=vis, ewa_WH=torch.tensor([W, H]),
             ewa_conic=conic.detach(), ewa_radius=radius, ewa_tiles=tiles, ewa_dxyz=x1.grad, ewa_dcov3d=c1.grad,
             ewa_dintr=i1.grad, ewa_dextr=e1.grad)
    # ---- SH bases, degrees 0..10 (test_compute_sh.py:162-325)
    m = _import_ref_test("test_compute_sh")
    torch.manual_seed(123)
    d = torch.randn(64, 3, dtype=torch.float64); d = d / d.norm(dim=1, keepdim=True)
    G["sh_dirs"] = d
    for deg in range(11):
        G[f"sh_bases_{deg}"] = m.eval_sh_bases((deg + 1) ** 2, d)
    # ---- alpha blending, per-pixel loop oracle (test_alpha_blending.py:6-63,112-134: 32x16, N=20, C=33, bg=1)
    m = _import_ref_test("test_alpha_blending")
    torch.manual_seed(121)
    w, h, bg, N = 32, 16, 1, 20
    uv = torch.rand(N, 2); uv[:, 0] *= w; uv[:, 1] *= h
    A = torch.randn(N, 2, 2); cv = torch.bmm(A, A.transpose(1, 2))
    conic = torch.stack([cv[:, 0, 0], cv[:, 0, 1], cv[:, 1, 1]], -1)
    depth = torch.rand(N, 1) * 5
    radius = (torch.rand(N, 1) * 5).int()
    tiles = m.get_tiles(uv, radius.squeeze(-1), w, h)
    opacity = torch.rand(N, 1)
    feature = torch.rand(N, 33)
    ids, tr = O.sort_gaussian(uv, depth, w, h, radius, tiles)
    u1, c1, o1, f1 = uv.clone().requires_grad_(), conic.clone().requires_grad_(), opacity.clone().requires_grad_(), feature.clone().requires_grad_()
    img = m.alpha_blending_torch_impl(u1, c1, o1, f1, ids, tr, bg, w, h)
    img.sum().backward()
    G.update(ab_uv=uv, ab_conic=conic, ab_depth=depth, ab_radius=radius, ab_tiles=tiles, ab_opacity=opacity, ab_feature=feature,
             ab_ids=ids, ab_tr=tr, ab_img=img.detach(), ab_duv=u1.grad, ab_dconic=c1.grad, ab_dop=o1.grad, ab_dfeat=f1.grad)
    np.savez_compressed(out_path, **{k: v.detach().cpu().numpy() for k, v in G.items()})
    print("wrote", out_path, {k: tuple(v.shape) for k, v in G.items()})


def from_ref_gpu(out_path):
    from oracle import ref_driver
    from pointrix_b200 import scene

    assert ref_driver.available(), "needs oracle/_ref and a GPU"
    P, W, H = 3000, 208, 120
    c, sc, cams = scene.make_config("cfg1", P=P, views=1)
    cams = scene.make_cameras(1, W, H, seed=1)
    scd = {k: v.cuda() for k, v in sc.items()}
    E, intr, cc = cams["extrinsic_matrix"][0].cuda(), cams["intrinsic_params"].cuda(), cams["camera_center"][0].cuda()
    g = torch.Generator().manual_seed(11)
    extra = torch.randn(P, 5, generator=g)
    G = dict(P=torch.tensor(P), W=torch.tensor(W), H=torch.tensor(H), E=E, intr=intr, cc=cc, extra=extra, **sc)
    for tag, deg, rd, ex in (("a", 3, False, None), ("b", 1, True, extra.cuda())):
        f = ref_driver.render_forward(H, W, E, intr, cc, **scd, sh_degree=deg, render_depth=rd, extra=ex)
        dimg = torch.randn(f["img"].shape, generator=g).cuda()
        b = ref_driver.render_backward(f, dimg, scd["position"], scd["opacity"], scd["scaling"], scd["rotation"], scd["shs"], cc,
                                       sh_degree=deg, render_depth=rd, n_extra=0 if ex is None else ex.shape[1], camera_grads=True)
        for k in ("uv", "depth", "cov3d", "conic", "radius", "tiles", "idx_sorted", "tile_range", "keys", "rgb", "img", "final_T", "ncontrib"):
            G[f"{tag}_{k}"] = f[k]
        G[f"{tag}_dimg"] = dimg
        for k, v in b.items():
            G[f"{tag}_g_{k}"] = v
    np.savez_compressed(out_path, **{k: v.detach().cpu().numpy() for k, v in G.items()})
    print("wrote", out_path)


def from_ref_loss(out_path):
    sys.modules.setdefault("lpips", types.ModuleType("lpips"))
    spec = importlib.util.spec_from_file_location("ref_loss", "/root/reference/pointrix/model/loss.py")
    L = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(L)
    G = {}
    g = torch.Generator().manual_seed(5)
    # (tag, shape): ragged sizes around the 32x32 tile and the 11-tap window, a 3-D image, smooth + noisy content
    cases = [("a", (2, 3, 37, 53)), ("b", (1, 3, 64, 96)), ("c", (3, 1, 9, 7)), ("d", (3, 40, 33)), ("e", (1, 4, 1, 1))]
    for tag, shp in cases:
        gt = torch.rand(shp, generator=g)
        if tag in ("a", "b"):  # a smooth ground truth and a prediction near it, like a half-trained render
            yy, xx = torch.meshgrid(torch.linspace(0, 3, shp[-2]), torch.linspace(0, 4, shp[-1]), indexing="ij")
            gt = (0.5 + 0.4 * torch.sin(xx + 2 * yy)).expand(shp).clone() + 0.05 * torch.rand(shp, generator=g)
            pred = (gt + 0.1 * torch.randn(shp, generator=g)).clamp(0, 1)
        else:
            pred = torch.rand(shp, generator=g)
        G[f"{tag}_pred"], G[f"{tag}_gt"] = pred, gt
        p1 = pred.clone().requires_grad_()
        l1 = L.l1_loss(p1, gt)
        l1.backward()
        G[f"{tag}_l1"], G[f"{tag}_l1_grad"] = l1.detach(), p1.grad
        p2 = pred.clone().requires_grad_()
        l2 = L.l2_loss(p2, gt)
        l2.backward()
        G[f"{tag}_l2"], G[f"{tag}_l2_grad"] = l2.detach(), p2.grad
        G[f"{tag}_l1_map"] = L.l1_loss(pred, gt, return_mean=False)
        if len(shp) == 4:
            G[f"{tag}_psnr"] = L.psnr(pred, gt)
            G[f"{tag}_ssim_per_image"] = L.ssim(pred, gt, size_average=False)
        p3 = pred.clone().requires_grad_()
        s = L.ssim(p3, gt)
        s.backward()
        G[f"{tag}_ssim"], G[f"{tag}_ssim_grad"] = s.detach(), p3.grad
        # BaseModel.get_loss_dict (base_model.py:117-120) with lambda_ssim = 0.2
        p4 = pred.clone().requires_grad_()
        loss = (1.0 - 0.2) * L.l1_loss(p4, gt) + 0.2 * (1.0 - L.ssim(p4, gt))
        loss.backward()
        G[f"{tag}_loss"], G[f"{tag}_loss_grad"] = loss.detach(), p4.grad
    G["window"] = L.create_window(11, 1)[0, 0]
    np.savez_compressed(out_path, **{k: v.detach().cpu().numpy() for k, v in G.items()})
    print("wrote", out_path, {k: tuple(v.shape) for k, v in G.items() if k.endswith("_pred")})


def load_reference_plugin():
    """The reference's OWN plugin file, pointrix/model/renderer/msplat.py, executed where it lies against stubs
    of what it imports: `BaseObject` (pointrix/utils/base.py:24-39 needs omegaconf -- a 15-line
    stand-in that parses the dataclass Config and calls setup), the REAL pointrix/utils/registry.py and
    pointrix/model/renderer/utils/renderer_utils.py loaded from the reference tree, and a module `msplat`
    whose six operators are the CPU oracle (oracle/msplat_oracle.py).  Returns the module: its MsplatRender
    class is the reference's code, line for line, driving our restatement of the operators."""
    import importlib.util
    import types
    from dataclasses import fields

    import torch

    from oracle import msplat_oracle as O

    ref = "/root/reference/pointrix"

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    for name in ("pointrix", "pointrix.utils", "pointrix.model", "pointrix.model.renderer", "pointrix.model.renderer.utils"):
        pkg(name)

    class BaseObject:  # pointrix/utils/base.py:24-39 without omegaconf
        def __init__(self, cfg=None, *args, **kwargs):
            known = {f.name for f in fields(self.Config)}
            self.cfg = self.Config(**{k: v for k, v in (cfg or {}).items() if k in known})
            self.setup(*args, **kwargs)

        def setup(self, *args, **kwargs):
            pass

    base = types.ModuleType("pointrix.utils.base")
    base.BaseObject = BaseObject
    sys.modules["pointrix.utils.base"] = base
    load("pointrix.utils.registry", os.path.join(ref, "utils", "registry.py"))
    load("pointrix.model.renderer.utils.renderer_utils", os.path.join(ref, "model", "renderer", "utils", "renderer_utils.py"))
    ms = types.ModuleType("msplat")
    for fn in ("compute_sh", "project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending"):
        setattr(ms, fn, getattr(O, fn))
    sys.modules["msplat"] = ms
    # the reference moves `position` to the GPU by hand (msplat.py:94-95); this container has none
    torch.Tensor.cuda = lambda self, *a, **k: self
    return load("pointrix.model.renderer.msplat", os.path.join(ref, "model", "renderer", "msplat.py"))


def from_ref_plugin(out_path):
    """Golden vectors of the plugin level: MsplatRender.render_iter / render_batch of the reference's own
    msplat.py (see load_reference_plugin) on a small seeded scene, with autograd gradients."""
    import numpy as np
    import torch

    from bench import load_scene_module

    scene = load_scene_module()
    mod = load_reference_plugin()
    P, W, H = 1500, 96, 64
    _, sc, _ = scene.make_config("cfg1", P=P, views=1)
    cams = scene.make_cameras(2, W, H, seed=1)
    g = torch.Generator().manual_seed(2)
    out = {"P": np.int64(P), "W": np.int64(W), "H": np.int64(H)}
    for k, v in sc.items():
        out[k] = v.numpy()
    for k, v in cams.items():
        out["cam_" + k] = v.numpy()
    for tag, deg, depth_ch in (("a", 3, False), ("b", 1, True)):
        r = mod.MsplatRender({"render_depth": depth_ch}, True, "cpu")
        r.sh_degree = deg
        leaves = {k: v.clone().requires_grad_() for k, v in sc.items()}
        o = r.render_iter(H, W, cams["extrinsic_matrix"][0], cams["intrinsic_params"], cams["camera_center"][0], **leaves)
        img = torch.cat(list(o["rendered_features_split"].values()), 0)
        dimg = torch.randn(img.shape, generator=g)
        img.backward(dimg)
        out[f"{tag}_img"], out[f"{tag}_dimg"] = img.detach().numpy(), dimg.numpy()
        out[f"{tag}_radii"] = o["radii"].numpy()
        out[f"{tag}_visibility"] = o["visibility"].numpy()
        out[f"{tag}_g_ndc"] = o["uv_points"].grad.numpy()
        for k, v in leaves.items():
            out[f"{tag}_g_{k}"] = v.grad.numpy()
    # render_batch: two views, reductions of msplat.py:160-213
    r = mod.MsplatRender({}, True, "cpu")
    r.sh_degree = 3
    with torch.no_grad():
        rb = r.render_batch(dict(height=H, width=W, extrinsic_matrix=cams["extrinsic_matrix"],
                                 intrinsic_params=cams["intrinsic_params"], camera_center=cams["camera_center"], **sc))
    out["batch_rgb"] = rb["rgb"].numpy()
    out["batch_radii"] = rb["radii"].numpy()
    out["batch_visibility"] = rb["visibility"].numpy()
    # update_sh_degree / state_dict (msplat.py:215-248)
    r2 = mod.MsplatRender({"update_sh_iter": 10, "max_sh_degree": 2}, False, "cpu")
    degs = []
    for step in range(0, 45):
        r2.update_sh_degree(step)
        degs.append(r2.sh_degree)
    out["sh_schedule"] = np.array(degs, dtype=np.int64)
    out["bg_black"] = np.float64(r2.bg_color)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith(("a_", "batch"))})


def from_ref_controller(out_path):
    """Golden vectors of DensificationController.preprocess (pointrix/controller/gs.py:259-333)."""
    import importlib.util
    import types

    import numpy as np
    import torch

    ref = "/root/reference/pointrix"

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    class BaseObject:
        pass

    for name in ("pointrix", "pointrix.utils", "pointrix.controller", "pointrix.model", "pointrix.model.utils"):
        mod(name)
    mod("pointrix.utils.config", C=lambda *a, **k: None)
    mod("pointrix.utils.base", BaseObject=BaseObject)
    mod("pointrix.utils.pose", quat_to_rotmat=lambda q: q)
    mod("pointrix.model.utils.gaussian_utils", sigmoid_inv=lambda x: x)
    load("pointrix.utils.registry", os.path.join(ref, "utils", "registry.py"))
    load("pointrix.controller.base", os.path.join(ref, "controller", "base.py"))
    gs = load("pointrix.controller.gs", os.path.join(ref, "controller", "gs.py"))
    ctl = object.__new__(gs.DensificationController)
    P, W, H = 4000, 640, 360
    ctl.cfg = types.SimpleNamespace(normalize_grad=True)
    ctl.width, ctl.height, ctl.device = W, H, "cpu"
    ctl.grad_accum = torch.zeros((P, 1))
    ctl.acc_steps = torch.zeros((P, 1))
    ctl.max_radii = torch.zeros((P,))
    g = torch.Generator().manual_seed(5)
    out = {"P": np.int64(P), "W": np.int64(W), "H": np.int64(H)}
    for it in range(3):
        views = []
        radii = torch.zeros(P, dtype=torch.int32)
        for v in range(2):
            uv = torch.zeros(P, 2, requires_grad=True)
            r_v = (torch.rand(P, generator=g) < 0.6) * torch.randint(1, 40, (P,), generator=g, dtype=torch.int32)
            uv.grad = torch.randn(P, 2, generator=g) * 1e-4 * (r_v > 0)[:, None]
            views.append(uv)
            radii = torch.maximum(radii, r_v.to(torch.int32))
        vis = radii > 0
        out[f"it{it}_uvgrad"] = torch.stack([u.grad for u in views]).numpy()
        out[f"it{it}_radii"] = radii.numpy()
        ctl.preprocess(uv_points=views, visibility=vis, radii=radii)
        out[f"it{it}_grad_accum"] = ctl.grad_accum.clone().numpy()
        out[f"it{it}_acc_steps"] = ctl.acc_steps.clone().numpy()
        out[f"it{it}_max_radii"] = ctl.max_radii.clone().numpy()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)


def from_ref_densify(out_path):
    """Golden vectors of DensificationController.densify (pointrix/controller/gs.py:72-234, 286-313, 336-341;
    base.py:89-96) with the real point-cloud surgery (points.py:179-357, point_utils.py:51-106)."""
    import importlib.util
    import types
    from dataclasses import fields

    import numpy as np
    import torch

    ref = "/root/reference/pointrix"

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    class BaseObject:  # pointrix/utils/base.py:24-39 without omegaconf
        def __init__(self, cfg=None, *args, **kwargs):
            super().__init__()
            known = {f.name for f in fields(self.Config)}
            self.cfg = self.Config(**{k: v for k, v in (cfg or {}).items() if k in known})
            self.setup(*args, **kwargs)

        def setup(self, *args, **kwargs):
            pass

    class BaseModule(BaseObject, torch.nn.Module):
        pass

    for name in ("pointrix", "pointrix.utils", "pointrix.controller", "pointrix.model", "pointrix.model.utils",
                 "pointrix.model.point_cloud", "pointrix.model.point_cloud.utils", "pointrix.logger"):
        mod(name)
    mod("pointrix.utils.config", C=lambda *a, **k: None)
    mod("pointrix.utils.base", BaseObject=BaseObject, BaseModule=BaseModule)
    mod("pointrix.logger.writer", Logger=types.SimpleNamespace(log=lambda *a, **k: None))
    mod("plyfile", PlyData=None, PlyElement=None)  # the un-vendored PLY codec: not touched by the surgery
    load("pointrix.utils.registry", os.path.join(ref, "utils", "registry.py"))
    pose = load("pointrix.utils.pose", os.path.join(ref, "utils", "pose.py"))
    pu = load("pointrix.model.point_cloud.utils.point_utils", os.path.join(ref, "model", "point_cloud", "utils", "point_utils.py"))
    mod("pointrix.model.utils.gaussian_utils", sigmoid_inv=pu.sigmoid_inv)
    load("pointrix.model.point_cloud.points", os.path.join(ref, "model", "point_cloud", "points.py"))
    gp = load("pointrix.model.point_cloud.gaussian_points", os.path.join(ref, "model", "point_cloud", "gaussian_points.py"))
    load("pointrix.controller.base", os.path.join(ref, "controller", "base.py"))
    gs = load("pointrix.controller.gs", os.path.join(ref, "controller", "gs.py"))

    # the split samples with torch.zeros(..., device="cuda") (gs.py:166): this container has no GPU
    zeros = torch.zeros
    torch.zeros = lambda *a, **k: zeros(*a, **{kk: ("cpu" if kk == "device" and str(vv).startswith("cuda") else vv) for kk, vv in k.items()})

    np.random.seed(0)
    torch.manual_seed(0)
    P0 = 800
    init = types.SimpleNamespace(init_type="random", num_points=P0, radius=1.0, feat_dim=3)
    pc = gp.GaussianPointCloud({"initializer": init, "max_sh_degree": 3})
    names = ("position", "features", "features_rest", "scaling", "rotation", "opacity")
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():  # a trained-looking table: spread scales, opacities on both sides of the pruning threshold
        pc.scaling.add_(torch.randn(P0, 3, generator=g) * 0.8)
        pc.rotation.copy_(torch.randn(P0, 4, generator=g))
        pc.opacity.copy_(torch.randn(P0, 1, generator=g) * 3.0 - 1.0)
        pc.features_rest.copy_(torch.randn(P0, 15, 3, generator=g) * 0.1)
    lrs = {"position": 0.00016, "features": 0.0025, "features_rest": 0.000125, "scaling": 0.005, "rotation": 0.001, "opacity": 0.05}
    opt = torch.optim.Adam([{"params": [getattr(pc, n)], "lr": lrs[n], "name": "point_cloud." + n} for n in names], eps=1e-15)
    for _ in range(3):  # populate the Adam state
        for n in names:
            getattr(pc, n).grad = torch.randn(getattr(pc, n).shape, generator=g) * 1e-3
        opt.step()

    ctl = object.__new__(gs.DensificationController)
    ctl.cfg = types.SimpleNamespace(normalize_grad=True, max_points=5000000, densify_stop_iter=15000, densify_start_iter=500)
    ctl.device, ctl.point_cloud, ctl.optimizer, ctl.cameras_extent = "cpu", pc, opt, 2.5
    ctl.split_num, ctl.min_opacity, ctl.prune_interval, ctl.densify_grad_threshold = 2, 0.005, 100, 0.0002
    ctl.duplicate_interval, ctl.opacity_reset_interval, ctl.percent_dense = 100, 3000, 0.01
    ctl.width, ctl.height = 640, 360

    def snapshot(tag, out):
        for n in names:
            p_ = getattr(pc, n)
            out[f"{tag}_{n}"] = p_.detach().clone().numpy()
            st = opt.state[p_]
            out[f"{tag}_{n}_exp_avg"] = st["exp_avg"].clone().numpy()
            out[f"{tag}_{n}_exp_avg_sq"] = st["exp_avg_sq"].clone().numpy()
            out[f"{tag}_{n}_step"] = np.float64(float(st["step"]))
        out[f"{tag}_grad_accum"] = ctl.grad_accum.clone().numpy()
        out[f"{tag}_acc_steps"] = ctl.acc_steps.clone().numpy()
        out[f"{tag}_max_radii"] = ctl.max_radii.clone().numpy()

    out = {"cameras_extent": np.float64(ctl.cameras_extent)}
    for case, step in enumerate((600, 3000, 3100)):
        n_pts = len(pc)
        ctl.step = step
        ctl.grad_accum = torch.rand(n_pts, 1, generator=g) * 6e-4          # average gradients around the 2e-4 threshold
        ctl.acc_steps = torch.randint(0, 4, (n_pts, 1), generator=g).float()  # zeros: 0/0 -> nan -> 0 in densify()
        ctl.max_radii = torch.rand(n_pts, generator=g) * 30.0
        snapshot(f"c{case}_before", out)
        out[f"c{case}_step"] = np.int64(step)
        torch.manual_seed(100 + case)                                     # the split's torch.normal draws from here
        orig_empty = torch.cuda.empty_cache
        torch.cuda.empty_cache = lambda: None
        try:
            ctl.densify()
        finally:
            torch.cuda.empty_cache = orig_empty
        snapshot(f"c{case}_after", out)
    torch.zeros = zeros
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items() if k.endswith("after_position")})


if __name__ == "__main__":
    if "--from-ref-densify" in sys.argv:
        from_ref_densify(os.path.join(ROOT, "tests", "golden", "ref_densify.npz"))
    elif "--from-ref-controller" in sys.argv:
        from_ref_controller(os.path.join(ROOT, "tests", "golden", "ref_controller.npz"))
    elif "--from-ref-plugin" in sys.argv:
        from_ref_plugin(os.path.join(ROOT, "tests", "golden", "ref_plugin.npz"))
    elif "--from-ref-loss" in sys.argv:
        from_ref_loss(os.path.join(ROOT, "tests", "golden", "ref_loss.npz"))
    elif "--from-ref-tests" in sys.argv:
        from_ref_tests(os.path.join(ROOT, "tests", "golden", "ref_test_oracles.npz"))
    elif "--from-ref-gpu" in sys.argv:
        out = sys.argv[-1] if sys.argv[-1].endswith(".npz") else os.path.join(ROOT, "gpurun_out", "ref_gpu_small.npz")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        from_ref_gpu(out)
    else:
        print(__doc__)
