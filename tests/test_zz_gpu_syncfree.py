"""Sync-free forward (MsplatRender.Config.sync_free): no host wait in a no_grad forward, the capacity check trails by one
call, and the forward can be captured into a CUDA graph and replayed with new cameras (the GUI's render loop,
pointrix/webgui/gui.py:160-209).  All three pass on a B200 (round 2, the last 13 seconds of the GPU budget).
"""
import pytest
import torch

from tests.util import scene_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import pointrix_b200

    return pointrix_b200


def _setup(pb, P=50_000, W=640, H=360, sync_free=True):
    c, sc, cams = scene_inputs("cfg2", P=P, views=3, W=W, H=H)
    r = pb.parse_renderer({"name": "MsplatRender", "sync_free": sync_free}, white_bg=True, device="cuda:0")
    r.sh_degree = 3
    return r, sc, cams, W, H


def test_sync_free_forward_equals_the_checked_forward(pb):
    from pointrix_b200 import renderer

    r_sf, sc, cams, W, H = _setup(pb)
    r_ck, *_ = _setup(pb, sync_free=False)
    with torch.no_grad():
        for v in range(3):
            a = r_ck.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **sc)
            b = r_sf.render_iter(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **sc)
            assert torch.equal(a["rendered_features_split"]["rgb"], b["rendered_features_split"]["rgb"])
            assert torch.equal(a["radii"], b["radii"])
    renderer.check_sync_free()
    # with autograd on, the flag is ignored (training keeps the exact, checked path)
    leaves = {k: v.clone().requires_grad_() for k, v in sc.items()}
    out = r_sf.render_iter(H, W, cams["extrinsic_matrix"][0], cams["intrinsic_params"], cams["camera_center"][0], **leaves)
    out["rendered_features_split"]["rgb"].sum().backward()
    assert leaves["position"].grad is not None


def test_sync_free_overflow_is_reported_by_the_next_call(pb):
    from pointrix_b200 import renderer

    r, sc, cams, W, H = _setup(pb)
    P = sc["position"].shape[0]
    with torch.no_grad():
        r.render_iter(H, W, cams["extrinsic_matrix"][0], cams["intrinsic_params"], cams["camera_center"][0], **sc)
        renderer.check_sync_free()
        renderer._CAPACITY[(0, P, W, H)] = 1024          # far too small on purpose
        r.render_iter(H, W, cams["extrinsic_matrix"][0], cams["intrinsic_params"], cams["camera_center"][0], **sc)
        with pytest.raises(RuntimeError, match="truncated"):
            renderer.check_sync_free()
        assert renderer._CAPACITY[(0, P, W, H)] > 1024   # raised for the re-render
        out = r.render_iter(H, W, cams["extrinsic_matrix"][0], cams["intrinsic_params"], cams["camera_center"][0], **sc)
        renderer.check_sync_free()
    r_ck, *_ = _setup(pb, sync_free=False)
    with torch.no_grad():
        ref = r_ck.render_iter(H, W, cams["extrinsic_matrix"][0], cams["intrinsic_params"], cams["camera_center"][0], **sc)
    assert torch.equal(out["rendered_features_split"]["rgb"], ref["rendered_features_split"]["rgb"])


def test_sync_free_forward_is_cuda_graph_capturable(pb):
    from pointrix_b200 import renderer

    r, sc, cams, W, H = _setup(pb)
    E = cams["extrinsic_matrix"][0].clone()
    cc = cams["camera_center"][0].clone()
    intr = cams["intrinsic_params"]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(3):  # warm-up on the capture stream: workspace, one-time initialisation, capacity
            r.render_iter(H, W, E, intr, cc, **sc)
        renderer.check_sync_free()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        out = r.render_iter(H, W, E, intr, cc, **sc)
    img = out["rendered_features_split"]["rgb"]
    r_ck, *_ = _setup(pb, sync_free=False)
    for v in (1, 2, 0):  # new camera written into the captured inputs, replay, compare with the checked path
        E.copy_(cams["extrinsic_matrix"][v])
        cc.copy_(cams["camera_center"][v])
        g.replay()
        torch.cuda.synchronize()
        with torch.no_grad():
            ref = r_ck.render_iter(H, W, cams["extrinsic_matrix"][v], intr, cams["camera_center"][v], **sc)
        assert torch.equal(img, ref["rendered_features_split"]["rgb"]), v
