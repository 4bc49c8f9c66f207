"""On-disk formats of the Gaussian table -- host code, no GPU involved.

``save_ply`` / ``load_ply`` write and read the file ``PointCloud.save_ply`` / ``load_ply`` produce
(``pointrix/model/point_cloud/points.py:359-427``): one ``vertex`` element, every property ``float``
(``'f4'``), columns ``x y z nx ny nz`` followed by ``{name}_{i}`` for every registered attribute but
``position`` in registration order with its trailing dimensions flattened row-major
(``list_of_attributes``, ``points.py:359-369``); for the Gaussian point cloud that order is ``features``
[P,1,3], ``features_rest`` [P,15,3], ``scaling`` [P,3], ``rotation`` [P,4], ``opacity`` [P,1]
(``points.py:52-62``, ``gaussian_points.py:28-51``), all PRE-activation, normals written as zeros.
The reference delegates the encoding to the third-party ``plyfile`` package (absent here, version not
pinned by the reference): this module restates the PLY 1.0 layout that package writes for such an
element -- ``format binary_little_endian 1.0``, ``property float <name>`` lines, rows packed without
padding.  PARITY UNPINNED against plyfile itself; pinned against the PLY 1.0 specification by
tests/test_io.py (hand-built files, ascii and both byte orders on the read side).

``save_checkpoint`` / ``load_checkpoint`` keep the keys of ``BaseTrainer.save_model`` / ``load_model``
(``pointrix/engine/base_trainer.py:145-171``) and of ``CheckPointHook.after_train``
(``pointrix/hook/checkpoint_hook.py:33-43``).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

GAUSSIAN_ATTRIBUTES: Tuple[Tuple[str, Tuple[int, ...]], ...] = (
    ("position", (3,)), ("features", (1, 3)), ("features_rest", (15, 3)), ("scaling", (3,)), ("rotation", (4,)),
    ("opacity", (1,)),
)

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
    "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
    "double": "f8", "float64": "f8",
}


def list_of_attributes(table: Mapping[str, torch.Tensor]) -> list:
    """Column names of the ply file for an ordered attribute table (``points.py:359-369``)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    for name, value in table.items():
        if name != "position":
            names += [f"{name}_{i}" for i in range(int(np.prod(value.shape[1:])))]
    return names


def save_ply(path, table: Mapping[str, torch.Tensor]) -> None:
    """Write the attribute table (``position`` first, registration order) like ``PointCloud.save_ply``."""
    if "position" not in table or next(iter(table)) != "position":
        raise ValueError("the table must start with 'position'")
    d = os.path.dirname(str(path))
    if d:
        os.makedirs(d, exist_ok=True)
    pos = table["position"].detach().cpu().numpy().astype(np.float32)
    n = pos.shape[0]
    cols = [pos.reshape(n, 3), np.zeros((n, 3), dtype=np.float32)]
    for name, value in table.items():
        if name != "position":
            if value.shape[0] != n:
                raise ValueError(f"attribute {name!r} has {value.shape[0]} rows, position has {n}")
            cols.append(value.detach().reshape(n, int(np.prod(value.shape[1:]))).cpu().numpy().astype(np.float32))
    data = np.ascontiguousarray(np.concatenate(cols, axis=1).astype("<f4"))
    names = list_of_attributes(table)
    assert data.shape[1] == len(names)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    header += [f"property float {c}" for c in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(data.tobytes())


def read_ply_vertices(path) -> "OrderedDict[str, np.ndarray]":
    """Columns of the first element of a PLY file (ascii, binary_little_endian or binary_big_endian;
    scalar properties only, which is all ``save_ply`` writes)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: header without end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported")
                if tok[1] not in _PLY_TYPES:
                    raise ValueError(f"{path}: unknown property type {tok[1]!r}")
                elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt is None or not elements:
            raise ValueError(f"{path}: incomplete header")
        name, count, props = elements[0]
        if fmt == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(props)))
            if rows.shape != (count, len(props)):
                raise ValueError(f"{path}: expected {count} x {len(props)} values, found {rows.shape}")
            return OrderedDict((p, rows[:, i].astype(t)) for i, (p, t) in enumerate(props))
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unknown format {fmt!r}")
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(p, order + t) for p, t in props])
        raw = f.read(count * dt.itemsize)
        if len(raw) != count * dt.itemsize:
            raise ValueError(f"{path}: truncated ({len(raw)} of {count * dt.itemsize} data bytes)")
        rec = np.frombuffer(raw, dtype=dt, count=count)
        return OrderedDict((p, rec[p].astype(rec[p].dtype.newbyteorder("="))) for p, _ in props)


def load_ply(path, shapes: Optional[Mapping[str, Sequence[int]]] = None, device="cpu") -> Dict[str, torch.Tensor]:
    """Read a table written by ``save_ply`` (or by the reference) like ``PointCloud.load_ply``
    (``points.py:397-427``): ``position`` from x/y/z, every other attribute from its ``{name}_{i}``
    columns reshaped to ``[-1, *shapes[name]]``.  ``shapes`` defaults to the Gaussian point cloud's."""
    cols = read_ply_vertices(path)
    shapes = dict(GAUSSIAN_ATTRIBUTES) if shapes is None else {k: tuple(v) for k, v in shapes.items()}
    out: Dict[str, torch.Tensor] = OrderedDict()
    out["position"] = torch.from_numpy(np.stack([cols["x"], cols["y"], cols["z"]], axis=1)).float()
    for name, shp in shapes.items():
        if name == "position":
            continue
        k = int(np.prod(shp))
        try:
            value = np.stack([np.asarray(cols[f"{name}_{i}"]) for i in range(k)], axis=1)
        except KeyError as e:
            raise KeyError(f"{path}: column {e.args[0]!r} of attribute {name!r} is missing") from None
        out[name] = torch.from_numpy(value.reshape(-1, *shp)).float()
    return {k: v.to(device) for k, v in out.items()}


def save_checkpoint(path, global_step: int, optimizer_state: dict, model_state: Optional[dict] = None,
                    point_cloud_state: Optional[dict] = None) -> None:
    """``{"global_step", "optimizer", "model"}`` (``BaseTrainer.save_model``) and/or ``"point_cloud"``
    (``CheckPointHook.after_train``)."""
    data = {"global_step": int(global_step), "optimizer": optimizer_state}
    if model_state is not None:
        data["model"] = model_state
    if point_cloud_state is not None:
        data["point_cloud"] = point_cloud_state
    d = os.path.dirname(str(path))
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save(data, path)


def load_checkpoint(path, map_location="cpu", weights_only: bool = True) -> dict:
    """The dict ``BaseTrainer.load_model`` iterates over (``base_trainer.py:145-160``).  Tensors, numbers,
    strings and containers only by default; ``weights_only=False`` (the reference's ``torch.load`` default
    at the time, which executes arbitrary pickled code) is an explicit opt-in for trusted files."""
    data = torch.load(path, map_location=map_location, weights_only=weights_only)
    if not isinstance(data, dict) or "global_step" not in data:
        raise ValueError(f"{path}: not a pointrix checkpoint")
    return data
