"""Host-side logic of the plugin boundary (no GPU): registry, config, SH warm-up, state dict,
feature combine/split, argument checking, view sharding and the gloo all-reduce step."""
import os

import pytest
import torch

import pointrix_b200 as pb
from pointrix_b200 import parallel


def _renderer(**cfg):
    return pb.parse_renderer({"name": "MsplatRender", **cfg}, white_bg=True, device="cuda:0")


def test_registry_and_config():
    assert "MsplatRender" in pb.RENDERER_REGISTRY
    r = _renderer()
    assert (r.cfg.update_sh_iter, r.cfg.max_sh_degree, r.cfg.render_depth) == (1000, 3, False)
    assert r.sh_degree == 0 and r.bg_color == 1.0
    assert pb.parse_renderer({"name": "MsplatRender"}, white_bg=False, device="cuda:0").bg_color == 0.0
    with pytest.raises(KeyError):
        _renderer(not_a_field=1)
    with pytest.raises(KeyError):
        pb.parse_renderer({"name": "NoSuchRender"}, white_bg=True, device="cuda:0")


def test_sh_warmup_and_state_dict():
    """pointrix/model/renderer/msplat.py:215-248"""
    r = _renderer(update_sh_iter=10, max_sh_degree=2)
    seen = []
    for step in range(0, 45):
        r.update_sh_degree(step)
        seen.append(r.sh_degree)
    assert seen[0] == 1 and seen[9] == 1 and seen[10] == 2 and seen[-1] == 2  # +1 at 0, 10; capped at max
    sd = r.state_dict()
    assert sd == {"sh_degree": 2}
    r2 = _renderer()
    r2.load_state_dict(sd)
    assert r2.sh_degree == 2


def test_render_features_combine_split():
    a, b = torch.rand(5, 3), torch.rand(5, 1)
    rf = pb.RenderFeatures(rgb=a, depth=b)
    comb = rf.combine()
    assert comb.shape == (5, 4) and torch.equal(comb[:, :3], a)
    img = torch.rand(4, 6, 7)
    sp = rf.split(img)
    assert list(sp) == ["rgb", "depth"] and sp["rgb"].shape == (3, 6, 7) and torch.equal(sp["depth"], img[3:4])


def test_cpu_tensors_are_rejected():
    """the reference's CHECK_INPUT: '<x> must be a CUDA tensor' (msplat/msplat/include/utils.h:9-10)"""
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pb.project_point(torch.zeros(4, 3), torch.zeros(4), torch.zeros(3, 4), 8, 8)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pb.compute_sh(torch.zeros(4, 3, 16), torch.zeros(4, 3))
    r = _renderer()
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        r.render_iter(8, 8, torch.eye(4), torch.zeros(4), torch.zeros(3), torch.zeros(4, 3), torch.zeros(4, 1),
                      torch.zeros(4, 3), torch.zeros(4, 4), torch.zeros(4, 16, 3))


def test_loss_module_mirrors_reference_and_rejects_cpu():
    """pointrix_b200.loss keeps the names of pointrix/model/loss.py and has no CPU path."""
    from pointrix_b200 import loss

    for name in ("l1_loss", "l2_loss", "psnr", "ssim", "gaussian", "create_window", "l1_ssim_loss", "get_loss_dict"):
        assert callable(getattr(loss, name))
    w = loss.create_window(11, 3)
    assert w.shape == (3, 1, 11, 11) and abs(float(w[0, 0].sum()) - 1.0) < 1e-6
    x = torch.rand(1, 3, 16, 16)
    for fn in (loss.l1_loss, loss.l2_loss, loss.ssim, loss.l1_ssim_loss, loss.psnr):
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            fn(x, x)


def test_no_oracle_on_the_product_path():
    """The package never imports oracle/ (CPU restatement) -- there is no fallback."""
    import pathlib

    root = pathlib.Path(pb.__file__).parent
    for f in root.rglob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f


def test_shard_views():
    for world in (1, 2, 4, 8):
        for step in (0, 3):
            got = sorted(v for r in range(world) for v in parallel.shard_views(16, r, world, step))
            assert got == list(range(16))
    assert parallel.shard_views(16, 1, 4) == [1, 5, 9, 13]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grads = [torch.randn(50, 3, generator=g), torch.randn(50, 16, 3, generator=g)]
    ndc = torch.randn(50, 2, generator=g)
    radii = torch.randint(0, 9, (50,), generator=g, dtype=torch.int32)
    keep = [x.clone() for x in grads] + [ndc.clone(), radii.clone()]
    vis = parallel.allreduce_step(grads, ndc, radii, world)
    q.put((rank, [x.numpy() for x in keep], [x.numpy() for x in grads] + [ndc.numpy(), radii.numpy(), vis.numpy()]))
    dist.destroy_process_group()


def test_allreduce_step_gloo_world2():
    """world_size ranks x 1 view == one reference batch of world_size views.  average=True (each rank
    back-propagated its un-scaled view loss): parameter grads become the mean AND every view's ndc grad
    carries 1/world before the sum -- the reference's mean loss scales each view's ndc.grad by 1/B before
    accumulate_viewspace_grad sums them (pointrix/model/loss.py:27-46, controller/gs.py:274-278); radii
    max-reduced, visibility = any."""
    import numpy as np
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, k0, o0), (_, k1, o1) = res
    for i in range(2):
        np.testing.assert_allclose(o0[i], (k0[i] + k1[i]) / 2, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(o1[i], o0[i])
    np.testing.assert_allclose(o0[2], (k0[2] + k1[2]) / 2, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o1[2], o0[2])
    assert (o0[3] == np.maximum(k0[3], k1[3])).all() and (o0[4] == (o0[3] > 0)).all()


def _worker_flat(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(200 + rank)
    P = 40
    flat = torch.randn(61 * P, generator=g)
    # the fused backward's layout: shs | rotation | position | scaling | opacity | ndc
    views = [flat[0:48 * P].view(P, 16, 3), flat[48 * P:52 * P].view(P, 4), flat[52 * P:55 * P].view(P, 3),
             flat[55 * P:58 * P].view(P, 3), flat[58 * P:59 * P].view(P, 1)]
    ndc = flat[59 * P:61 * P].view(P, 2)
    assert len(parallel._coalesce(views + [ndc])) == 1  # one collective for the whole set
    radii = torch.randint(0, 9, (P,), generator=g, dtype=torch.int32)
    keep = flat.clone()
    rw = parallel.begin_radii_reduce(radii, world)
    parallel.allreduce_step(views, ndc, radii, world, average=False, radii_work=rw)
    q.put((rank, keep.numpy(), flat.numpy(), radii.numpy()))
    dist.destroy_process_group()


def test_allreduce_step_flat_buffer_sum_gloo_world2():
    """average=False (the loss carries 1/world): the flat gradient buffer of the fused backward goes
    out as ONE summed all-reduce; radii reduced by the handle started after the forward."""
    import numpy as np
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker_flat, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, k0, o0, r0), (_, k1, o1, r1) = res
    np.testing.assert_allclose(o0, k0 + k1, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o1, o0)
    assert (r0 == r1).all()


def test_coalesce_leaves_unrelated_tensors_alone():
    a, b = torch.zeros(4, 3), torch.zeros(5)
    out = parallel._coalesce([a, b])
    assert len(out) == 2
    flat = torch.zeros(10)
    gap = parallel._coalesce([flat[0:4], flat[6:10]])  # not contiguous: two collectives
    assert len(gap) == 2


def _worker_factored(rank, world, port, q):
    """The SH-factored exchange protocol (parallel.ShFactoredExchange) restated with gloo collectives and the
    CPU oracle: all-reduce 13 floats + all-gather (d_rgb, camera centre) + local rebuild == all-reduce of all 61."""
    import torch.distributed as dist

    from oracle import msplat_oracle as O
    from pointrix_b200 import scene

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, W, H, deg = 300, 64, 48, 3
    c, sc, _ = scene.make_config("cfg1", P=P, views=1)
    cams = scene.make_cameras(world, W, H, seed=1)
    leaves = {k: v.clone().requires_grad_() for k, v in sc.items()}
    o = O.render_iter(H, W, cams["extrinsic_matrix"][rank], cams["intrinsic_params"], cams["camera_center"][rank],
                      **leaves, sh_degree=deg)
    dimg = scene.upstream_gradient(3, H, W, seed=2) / world
    (o["rendered_features_split"]["rgb"] * dimg).sum().backward()
    # what the 61-float all-reduce leaves on every rank
    full = {k: v.grad.clone() for k, v in leaves.items()}
    for t in full.values():
        dist.all_reduce(t)
    # factored: the SH gradient never travels.  The gated dL/drgb of this view is its DC row / C0.
    d_rgb = leaves["shs"].grad[:, 0, :] / 0.28209479177387814
    rgbs = [torch.empty_like(d_rgb) for _ in range(world)]
    cens = [torch.empty(3) for _ in range(world)]
    dist.all_gather(rgbs, d_rgb.contiguous())
    dist.all_gather(cens, cams["camera_center"][rank].contiguous())
    rebuilt = torch.zeros(P, 16, 3)
    for r_, c_ in zip(rgbs, cens):  # rank order: identical bits on every rank
        d = sc["position"] - c_.reshape(1, 3)
        d = d / d.norm(dim=1, keepdim=True)
        rebuilt += O.sh_bases(16, d)[:, :16, None] * r_[:, None, :]
    err = ((rebuilt - full["shs"]).norm() / full["shs"].norm()).item()
    chk = rebuilt.double().sum().reshape(1)
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    q.put((rank, err, bool(all(torch.equal(a, allc[0]) for a in allc)), float(full["shs"].abs().sum())))
    dist.destroy_process_group()


def test_factored_exchange_protocol_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker_factored, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same_bits, mag in res:
        assert mag > 0 and err <= 1e-5 and same_bits, (rank, err, same_bits)


def test_plugin_host_logic_matches_reference_plugin_golden():
    """update_sh_degree schedule and background colour of the reference's own MsplatRender
    (tests/golden/ref_plugin.npz, oracle/make_golden.py --from-ref-plugin) -- host logic, no GPU."""
    import numpy as np

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_plugin.npz"))
    r = pb.MsplatRender({"update_sh_iter": 10, "max_sh_degree": 2}, False, "cpu")
    degs = []
    for step in range(0, 45):
        r.update_sh_degree(step)
        degs.append(r.sh_degree)
    assert degs == z["sh_schedule"].tolist()
    assert r.bg_color == float(z["bg_black"])


def test_densification_surgery_matches_reference_controller_golden():
    """pointrix_b200.densify.DensificationController (clone / split / prune / opacity reset + Adam-state surgery) against
    tests/golden/ref_densify.npz: the reference's OWN GaussianPointCloud + DensificationController.densify executed
    where they lie on a populated torch.optim.Adam (oracle/make_golden.py --from-ref-densify), three schedule points."""
    import types

    import numpy as np

    from pointrix_b200 import densify, optim

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_densify.npz"))
    names = ("position", "features", "features_rest", "scaling", "rotation", "opacity")
    for case in range(3):
        pre, post = f"c{case}_before", f"c{case}_after"
        t = lambda k: torch.from_numpy(z[k])  # noqa: E731
        opt = types.SimpleNamespace(
            params={n: t(f"{pre}_{n}").clone().requires_grad_() for n in names},
            state={n: {"exp_avg": t(f"{pre}_{n}_exp_avg").clone(), "exp_avg_sq": t(f"{pre}_{n}_exp_avg_sq").clone()} for n in names})
        P = opt.params["position"].shape[0]
        stats = optim.DensificationStats(P, "cpu", 640, 360)
        stats.grad_accum, stats.acc_steps, stats.max_radii = t(f"{pre}_grad_accum").clone(), t(f"{pre}_acc_steps").clone(), t(f"{pre}_max_radii").clone()
        ctl = densify.DensificationController(opt, stats, densify.DensifyConfig(), cameras_extent=float(z["cameras_extent"]))
        ctl.step = int(z[f"c{case}_step"])
        torch.manual_seed(100 + case)  # the generator's seed for this case: the split's torch.normal draws from it
        ctl.densify()
        assert len(ctl) == z[f"{post}_position"].shape[0] != P, case
        for n in names:
            assert torch.allclose(opt.params[n].detach(), t(f"{post}_{n}"), rtol=1e-6, atol=1e-7), (case, n)
            assert opt.params[n].requires_grad and opt.params[n].is_leaf
            for key in ("exp_avg", "exp_avg_sq"):
                assert torch.equal(opt.state[n][key], t(f"{post}_{n}_{key}")), (case, n, key)
        assert torch.equal(stats.grad_accum, t(f"{post}_grad_accum")) and torch.equal(stats.acc_steps, t(f"{post}_acc_steps"))
        assert torch.equal(stats.max_radii, t(f"{post}_max_radii"))
    # schedule: nothing happens before densify_start_iter or off the intervals; the step counter advances
    ctl.step = 10
    assert ctl.f_step() is False and ctl.step == 11
    ctl.step = 601
    assert ctl.f_step() is False
