"""GPU parity of the fused optimizer step + densification statistics (csrc/optim.cu) through the
public Python API (= the C ABI) against
  (a) torch.optim.Adam on CPU -- the reference's optimizer IS torch.optim.Adam (pointrix/optimizer/optimizer.py:107-140)
      with the groups / learning rates / eps of examples/gaussian_splatting/configs/nerf.yaml:49-69;
  (b) tests/golden/ref_controller.npz: outputs of the reference's own DensificationController.preprocess.
Tolerances: parameters and moments relative 1e-6 (fp32, same formula, different FMA contraction); statistics 1e-6.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def OPT():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from pointrix_b200 import optim

    return optim


def _table(P, seed=0, split_shs=True):
    g = torch.Generator().manual_seed(seed)
    t = {"position": torch.randn(P, 3, generator=g), "scaling": torch.randn(P, 3, generator=g),
         "rotation": torch.randn(P, 4, generator=g), "opacity": torch.randn(P, 1, generator=g)}
    if split_shs:
        t["features"] = torch.randn(P, 1, 3, generator=g)
        t["features_rest"] = torch.randn(P, 15, 3, generator=g) * 0.1
    else:
        t["shs"] = torch.randn(P, 16, 3, generator=g)
    return t


def _close(a, b, tol=2e-6):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) <= tol


@pytest.mark.parametrize("P", [1, 1000, 300_001])
def test_adam_matches_torch_optim_adam(OPT, P):
    from oracle import optim_oracle as OO

    cpu = {k: v.clone().requires_grad_() for k, v in _table(P).items()}
    gpu = {k: v.clone().cuda().requires_grad_() for k, v in _table(P).items()}
    ref = OO.make_adam(cpu)
    ours = OPT.GaussianAdam(gpu)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        for k in cpu:
            gr = torch.randn(cpu[k].shape, generator=g) * (10.0 ** -(step % 3))
            if step == 2 and k == "opacity":
                gr = None  # a group without a gradient is skipped, as torch.optim does
            cpu[k].grad = gr
            gpu[k].grad = None if gr is None else gr.cuda()
        ref.step()
        ours.step()
        if step == 1:
            ours.lrs["position"] = ref.param_groups[0]["lr"] = 0.5 * ours.lrs["position"]  # a scheduler at work
    for i, k in enumerate(cpu):
        assert _close(gpu[k], cpu[k]), k
        st = ref.state[cpu[k]]
        assert _close(ours.state[k]["exp_avg"], st["exp_avg"]) and _close(ours.state[k]["exp_avg_sq"], st["exp_avg_sq"]), k
    # torch.optim.Adam's state layout round-trips
    sd = ours.state_dict()
    again = OPT.GaussianAdam({k: v.detach().clone().requires_grad_() for k, v in gpu.items()})
    again.load_state_dict(sd)
    assert again.steps == ours.steps and ours.steps["opacity"] == 3 and ours.steps["position"] == 4
    assert torch.equal(again.state["scaling"]["exp_avg"], ours.state["scaling"]["exp_avg"])


def test_single_shs_leaf_trains_as_the_reference_two_groups(OPT):
    """A model holding ONE shs[P,16,3] leaf: its DC row uses the `features` learning rate, rows 1..15 `features_rest`'s."""
    from oracle import optim_oracle as OO

    P = 5000
    split, one = _table(P, split_shs=True), _table(P, split_shs=False)
    one["shs"] = torch.cat([split["features"], split["features_rest"]], 1)
    cpu = {k: v.clone().requires_grad_() for k, v in split.items()}
    gpu = {k: v.clone().cuda().requires_grad_() for k, v in one.items()}
    ref, ours = OO.make_adam(cpu), OPT.GaussianAdam(gpu)
    g = torch.Generator().manual_seed(3)
    for step in range(3):
        gs = torch.randn(P, 16, 3, generator=g)
        cpu["features"].grad, cpu["features_rest"].grad = gs[:, :1].clone(), gs[:, 1:].clone()
        gpu["shs"].grad = gs.cuda()
        for k in ("position", "scaling", "rotation", "opacity"):
            gr = torch.randn(cpu[k].shape, generator=g)
            cpu[k].grad, gpu[k].grad = gr, gr.cuda()
        ref.step()
        ours.update_model()
        assert gpu["shs"].grad is None  # update_model: step + zero_grad(set_to_none=True)
    assert _close(gpu["shs"][:, :1], cpu["features"]) and _close(gpu["shs"][:, 1:], cpu["features_rest"])


def test_statistics_match_reference_controller_golden_and_fold_into_the_adam_launch(OPT):
    from pointrix_b200 import _lib

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_controller.npz"))
    P, W, H = int(z["P"]), int(z["W"]), int(z["H"])
    stats = OPT.DensificationStats(P, "cuda", W, H)
    params = {"position": torch.zeros(P, 3, device="cuda", requires_grad=True)}
    opt = OPT.GaussianAdam(params)
    for it in range(3):
        uvg = torch.from_numpy(z[f"it{it}_uvgrad"]).cuda()
        radii = torch.from_numpy(z[f"it{it}_radii"]).cuda()
        views = []
        for v in range(2):
            u = torch.zeros(P, 2, device="cuda", requires_grad=True)
            u.grad = uvg[v].clone()
            views.append(u)
        if it == 1:   # statistics alone (DensificationController.preprocess)
            stats.preprocess(views, radii > 0, radii)
        else:         # folded into the optimizer's launch
            params["position"].grad = torch.ones(P, 3, device="cuda")
            n0 = _lib.launch_count
            opt.step(stats=stats, uv_points=views, visibility=radii > 0, radii=radii)
            assert _lib.launch_count - n0 == 1
        for name in ("grad_accum", "acc_steps", "max_radii"):
            want = torch.from_numpy(z[f"it{it}_{name}"])
            assert _close(getattr(stats, name), want, 1e-6), (it, name)
    assert opt.steps["position"] == 2 and float(params["position"].abs().max()) > 0
    avg = stats.average_grad()
    assert torch.isfinite(avg).all() and avg.shape == (P, 1)


def test_argument_errors(OPT):
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        OPT.GaussianAdam({"position": torch.zeros(4, 3, requires_grad=True)})
    with pytest.raises(KeyError):
        OPT.GaussianAdam({"mystery": torch.zeros(4, 3, device="cuda", requires_grad=True)})
    s = OPT.DensificationStats(10, "cuda", 64, 64)
    with pytest.raises(RuntimeError, match="statistics hold"):
        s.preprocess(torch.zeros(7, 2, device="cuda"), None, torch.zeros(7, dtype=torch.int32, device="cuda"))
