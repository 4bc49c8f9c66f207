"""CPU tests of the on-disk formats: the PLY layout against the PLY 1.0 specification
(hand-built expected bytes), round trips at the Gaussian table's shapes, the reader on ascii / big-endian
files, error handling; checkpoint keys of base_trainer.py:145-171."""
import struct

import numpy as np
import pytest
import torch

from pointrix_b200 import io as pio


def _table(P, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {name: torch.randn(P, *shp, generator=g) for name, shp in pio.GAUSSIAN_ATTRIBUTES}


def test_attribute_names_follow_registration_order():
    names = pio.list_of_attributes(_table(2))
    assert names[:6] == ["x", "y", "z", "nx", "ny", "nz"]
    assert names[6:9] == ["features_0", "features_1", "features_2"]
    assert names[9] == "features_rest_0" and names[9 + 44] == "features_rest_44"
    assert names[-8:] == ["scaling_0", "scaling_1", "scaling_2", "rotation_0", "rotation_1", "rotation_2", "rotation_3",
                          "opacity_0"]
    assert len(names) == 6 + 3 + 45 + 3 + 4 + 1


def test_ply_bytes_match_the_specification(tmp_path):
    """2 points, 2 attributes: the exact file a PLY 1.0 writer produces for an all-float vertex element."""
    t = {"position": torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]), "opacity": torch.tensor([[0.5], [-1.5]])}
    p = tmp_path / "a.ply"
    pio.save_ply(p, t)
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\n"
              "property float z\nproperty float nx\nproperty float ny\nproperty float nz\nproperty float opacity_0\n"
              "end_header\n").encode("ascii")
    body = struct.pack("<7f", 1, 2, 3, 0, 0, 0, 0.5) + struct.pack("<7f", 4, 5, 6, 0, 0, 0, -1.5)
    assert p.read_bytes() == header + body


@pytest.mark.parametrize("P", [0, 1, 1000])
def test_ply_round_trip_is_bit_exact(tmp_path, P):
    t = _table(P, seed=P)
    p = tmp_path / "sub" / "pc.ply"  # the directory is created, like os.makedirs in points.py:380
    pio.save_ply(p, t)
    back = pio.load_ply(p)
    assert list(back) == [n for n, _ in pio.GAUSSIAN_ATTRIBUTES]
    for name, shp in pio.GAUSSIAN_ATTRIBUTES:
        assert back[name].shape == (P, *shp)
        assert torch.equal(back[name], t[name]), name
    cols = pio.read_ply_vertices(p)
    assert all(float(np.abs(cols[k]).sum()) == 0.0 for k in ("nx", "ny", "nz"))
    # row-major flattening of [P,15,3]: column 3k+c is coefficient k, channel c
    if P:
        assert np.array_equal(cols["features_rest_7"], t["features_rest"][:, 2, 1].numpy())


def test_reader_handles_ascii_big_endian_and_comments(tmp_path):
    a = tmp_path / "a.ply"
    a.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\nproperty float x\nproperty float y\n"
                 "property float z\nproperty uchar red\nend_header\n0 1 2 255\n3 4 5 7\n")
    c = pio.read_ply_vertices(a)
    assert c["x"].tolist() == [0.0, 3.0] and c["red"].tolist() == [255, 7] and c["red"].dtype == np.uint8
    b = tmp_path / "b.ply"
    b.write_bytes(b"ply\nformat binary_big_endian 1.0\nelement vertex 1\nproperty double x\nproperty int y\nend_header\n"
                  + struct.pack(">di", 2.5, -3))
    c = pio.read_ply_vertices(b)
    assert c["x"].tolist() == [2.5] and c["y"].tolist() == [-3]
    # custom shapes (a point cloud with other registered attributes)
    t = {"position": torch.zeros(3, 3), "rgb": torch.arange(9.0).reshape(3, 3)}
    p = tmp_path / "c.ply"
    pio.save_ply(p, t)
    back = pio.load_ply(p, shapes={"position": (3,), "rgb": (3,)})
    assert torch.equal(back["rgb"], t["rgb"])


def test_ply_errors(tmp_path):
    with pytest.raises(ValueError, match="start with 'position'"):
        pio.save_ply(tmp_path / "x.ply", {"opacity": torch.zeros(1, 1), "position": torch.zeros(1, 3)})
    with pytest.raises(ValueError, match="rows"):
        pio.save_ply(tmp_path / "x.ply", {"position": torch.zeros(2, 3), "opacity": torch.zeros(1, 1)})
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"plx\n")
    with pytest.raises(ValueError, match="not a PLY"):
        pio.read_ply_vertices(bad)
    p = tmp_path / "t.ply"
    pio.save_ply(p, _table(4))
    p.write_bytes(p.read_bytes()[:-5])
    with pytest.raises(ValueError, match="truncated"):
        pio.read_ply_vertices(p)
    q = tmp_path / "q.ply"
    pio.save_ply(q, {"position": torch.zeros(1, 3)})
    with pytest.raises(KeyError, match="features_0"):
        pio.load_ply(q)


def test_checkpoint_keys_and_round_trip(tmp_path):
    w = torch.nn.Parameter(torch.randn(5, 3))
    opt = torch.optim.Adam([{"params": [w], "name": "point_cloud.position", "lr": 1e-3}])
    w.grad = torch.ones_like(w)
    opt.step()
    p = tmp_path / "out" / "chkpnt00007.pth"
    pio.save_checkpoint(p, 7, opt.state_dict(), model_state={"point_cloud.position": w.detach()},
                        point_cloud_state={"position": w.detach()})
    d = pio.load_checkpoint(p)
    assert set(d) == {"global_step", "optimizer", "model", "point_cloud"} and d["global_step"] == 7
    assert d["optimizer"]["param_groups"][0]["name"] == "point_cloud.position"
    assert torch.equal(d["model"]["point_cloud.position"], w.detach())
    opt2 = torch.optim.Adam([{"params": [torch.nn.Parameter(torch.zeros(5, 3))], "name": "point_cloud.position"}])
    opt2.load_state_dict(d["optimizer"])
    assert torch.equal(opt2.state_dict()["state"][0]["exp_avg"], opt.state_dict()["state"][0]["exp_avg"])
    torch.save([1, 2], tmp_path / "no.pth")
    with pytest.raises(ValueError, match="not a pointrix checkpoint"):
        pio.load_checkpoint(tmp_path / "no.pth")
