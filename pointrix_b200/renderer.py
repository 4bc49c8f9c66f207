"""``MsplatRender`` -- the pointrix renderer plugin, B200-native.

Drop-in for ``pointrix/model/renderer/msplat.py:12-248``: same registry entry
name, ``Config``, ``setup``, ``render_iter`` / ``render_batch`` signatures and
returned dict (``rendered_features_split`` / per-feature stacks, ``uv_points``,
``visibility``, ``radii``), ``update_sh_degree``, ``state_dict`` /
``load_state_dict``.  One view is ONE foreign call forward (19 kernels queued back to
back from native code: the fused per-Gaussian stage, 9 for the depth order of the
visible Gaussians, 8 for key emission + tile sort + ranges, the blend) and one
backward (blend backward, fused per-Gaussian backward) instead of the reference's
~14 kernels + ~40 torch ops; the autograd graph of ``render_iter`` (SURVEY.md section 8a
"gradient routing") is reproduced by a single ``torch.autograd.Function``.

Beyond the reference's surface: ``render_iter_raw`` (the point cloud's raw parameters,
activations inside the kernels), ``camera_extrinsics`` (the camera model of one view),
``Config.sync_free`` (no host wait in ``no_grad`` forwards; CUDA-graph capturable).
"""
from __future__ import annotations

import collections
import ctypes as C
import time
from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib, ops
from ._lib import lib
from .ops import _f32, _p, _stream
from .registry import BaseObject, register_renderer


class RenderFeatures:
    """pointrix/model/renderer/utils/renderer_utils.py:4-71 -- per-Gaussian feature
    tensors combined on the last dim in kwarg order, split back by channel count."""

    def __init__(self, **kwargs) -> None:
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device) -> None:
        for k, v in self.__dict__.items():
            if isinstance(v, Tensor):
                setattr(self, k, v.to(device))

    def items(self):
        return [(k, v) for k, v in self.__dict__.items() if isinstance(v, Tensor)]

    def combine(self) -> Tensor:
        return torch.cat([v for _, v in self.items()], dim=-1)

    def split(self, feature: Tensor) -> Dict[str, Tensor]:
        out, start = {}, 0
        for k, v in self.items():
            end = start + v.shape[-1]
            out[k] = feature[start:end, ...]
            start = end
        return out


# Optional owner of the flat gradient buffer (parallel.NvlsGradExchange: symmetric memory, so the
# data-parallel exchange reduces the gradients in place over NVLink without a staging copy).
_GRAD_SINK = None


def set_grad_sink(sink) -> None:
    """``sink.next_buffer(numel) -> Tensor | None`` supplies the buffer the fused backward writes the
    P-sized gradients into (None: allocate).  A sink with ``plan(P, device, cam_center, sh_degree)`` returns
    ``(d_shs, d_rot, d_pos, d_scale, d_opacity, d_ndc, d_rgb)`` instead (or None to decline); with a
    non-None ``d_rgb[P,3]`` the backward writes the factored SH gradient there and leaves ``d_shs`` to
    the sink's exchange.  Pass None to detach."""
    global _GRAD_SINK
    _GRAD_SINK = sink


# per-device pinned word the forward's scan kernel stores the intersection count into, scratch
# workspaces per (device, stream, P, capacity, W, H), and the learnt intersection capacity
_COUNT_WORD = {}
_WORKSPACE = {}
_CAPACITY = {}
_raw_stream = torch._C._cuda_getCurrentRawStream


def _count_word(dev_index: int):
    w = _COUNT_WORD.get(dev_index)
    if w is None:
        t = torch.zeros(16, dtype=torch.int32).pin_memory()
        w = _COUNT_WORD[dev_index] = (t, C.c_int.from_address(t.data_ptr()))
    return w


# ---- sync-free forward (inference) ---------------------------------------------------------------------------------
# Normally every forward spins on the pinned word until the key emission has stored the intersection count: a view
# that overflows the learnt capacity is re-binned before anyone sees its image (exactness), at the price of one
# host <-> device rendezvous per view.  With MsplatRender.Config.sync_free (and only under torch.no_grad()) the check
# TRAILS: the call returns at once, the count lands in one of 15 pinned slots, and the next call (or
# check_sync_free()) reads it -- raising if that earlier image was truncated, after raising the capacity.  Nothing
# in the forward then waits on the device, which is what CUDA-graph capture needs (the GUI's render loop,
# pointrix/webgui/gui.py:160-209).
_SYNC_FREE = False
_PENDING: Dict[int, "collections.deque"] = {}
_NEXT_SLOT: Dict[int, int] = {}


def _slot_word(dev_index: int, slot: int):
    t, _ = _count_word(dev_index)
    return C.c_int.from_address(t.data_ptr() + 4 * slot)


def _check_pending(dev_index: int, block: bool) -> None:
    dq = _PENDING.get(dev_index)
    while dq:
        slot, cap, ckey = dq[0]
        w = _slot_word(dev_index, slot)
        n = w.value
        if n == -1:
            if not block:
                return
            n = _wait_count(w)
        dq.popleft()
        ops.LAST_N[(dev_index, True)] = n
        if n > cap:
            _CAPACITY[ckey] = int(n * 1.25) + 65536
            raise RuntimeError(f"pointrix_b200: a sync-free render binned {n} tile intersections into a capacity of {cap}: "
                               "that image was truncated.  The capacity has been raised; render the view again")


def check_sync_free(device=None) -> None:
    """Wait for the intersection counts of all sync-free renders still in flight on ``device`` (default: the current
    one) and raise if any of them overflowed its capacity."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    _check_pending(idx, block=True)


def _wait_count(word) -> int:
    """Spin until the device has stored the count (it does so ~0.2 ms after the call, while the
    blend kernel is queued behind it)."""
    n = word.value
    if n != -1:
        return n
    deadline = time.monotonic() + 30.0
    while True:
        for _ in range(4096):
            n = word.value
            if n != -1:
                return n
        if time.monotonic() > deadline:
            torch.cuda.synchronize()  # surfaces a launch failure as a CUDA error
            if word.value == -1:
                raise RuntimeError("pointrix_b200: the binning kernels never reported an intersection count")


def _stage_events(timer, names):
    """Timing events around the stages `names` (consecutive stages of one foreign call) the timer wants:
    (events or None per boundary, the void*[len(names)+1] the native code records, NULL = skip)."""
    if timer is None:
        return None, None
    n = len(names) + 1
    need = [False] * n
    for k, name in enumerate(names):
        if timer.wants(name):
            need[k] = need[k + 1] = True
    if not any(need):
        return None, None
    evs = [torch.cuda.Event(enable_timing=True) if f else None for f in need]
    for e in evs:
        if e is not None:
            e.record()  # torch creates the CUDA event on the first record
    return evs, (C.c_void_p * n)(*[e.cuda_event if e is not None else None for e in evs])


_FWD_STAGES = ("pxb_fused_forward", "pxb_bin_prepare", "pxb_sort_gaussian", "pxb_blend_forward")
_BWD_STAGES = ("pxb_blend_backward", "pxb_fused_backward")


class _FusedRender(torch.autograd.Function):
    """render_iter's differentiable core: (Gaussians, camera) -> blended features.
    One foreign call forward (pxb_render_forward), one backward (pxb_render_backward)."""

    @staticmethod
    def forward(ctx, position, opacity, scaling, rotation, shs, shs_rest, extra, intr, extr, cam_center, ndc, sh_degree,
                W, H, bg, with_depth, nearest):
        # shs_rest is None: post-activation inputs (render_iter).  Otherwise RAW mode (render_iter_raw): opacity /
        # scaling / rotation are the point cloud's raw parameters, shs = features[P,1,3], shs_rest = features_rest
        pos, op = _f32(position, "position"), _f32(opacity, "opacity")
        sc, rot, sh = _f32(scaling, "scaling"), _f32(rotation, "rotation"), _f32(shs, "shs")
        sh_r = _f32(shs_rest, "features_rest") if shs_rest is not None else None
        intr_c, extr_c, cc = _f32(intr, "intrinsic_params"), _f32(extr, "extrinsic_matrix"), _f32(cam_center, "camera_center")
        ex = _f32(extra, "extra features") if extra is not None else None
        dev = pos.device
        P = pos.shape[0]
        n_extra = 0 if ex is None else ex.shape[1]
        Cc = 3 + int(with_depth) + n_extra
        S = lib.pxb_record_stride(Cc)
        W, H = int(W), int(H)
        n_tiles = ((W + 15) // 16) * ((H + 15) // 16)
        E = torch.empty
        f32, i32 = torch.float32, torch.int32
        rec = E((P, S), dtype=f32, device=dev)
        depth = E((P,), dtype=f32, device=dev)
        radius = E((P,), dtype=i32, device=dev)
        tile_range = E((n_tiles, 2), dtype=i32, device=dev)
        out = E((Cc, H, W), dtype=f32, device=dev)
        final_T = E((H, W), dtype=f32, device=dev)
        ncontrib = E((H, W), dtype=i32, device=dev)
        _lib.ensure_init()
        host_t, word = _count_word(dev.index)
        ckey = (dev.index, P, W, H)
        cap = _CAPACITY.get(ckey) or max(4 * P, 1 << 20)
        timer = _lib._timer
        same_dev = torch.cuda.current_device() == dev.index
        while True:
            stream = _raw_stream(dev.index)
            wkey = (dev.index, stream, P, cap, W, H)
            ws = _WORKSPACE.get(wkey)
            if ws is None:
                for k in [k for k in _WORKSPACE if k[:2] == wkey[:2]]:
                    del _WORKSPACE[k]  # one live workspace per (device, stream)
                ws = _WORKSPACE[wkey] = E((lib.pxb_render_workspace_bytes(P, cap, W, H),), dtype=torch.uint8, device=dev)
            idx_sorted = E((cap,), dtype=i32, device=dev)
            evs, ev_arr = _stage_events(timer, _FWD_STAGES)
            sync_free = _SYNC_FREE and not torch.is_grad_enabled()
            slot = 0
            if sync_free:  # trailing check: earlier calls first, then a pinned slot of this call's own
                dq = _PENDING.setdefault(dev.index, collections.deque())
                _check_pending(dev.index, block=len(dq) >= 14)
                slot = _NEXT_SLOT[dev.index] = _NEXT_SLOT.get(dev.index, 0) % 15 + 1
                word = _slot_word(dev.index, slot)
            word.value = -1
            args = (P, int(sh_degree), _p(pos), _p(sc), _p(rot), _p(op), _p(sh), _p(sh_r), _p(ex), n_extra, int(with_depth),
                    _p(intr_c), _p(extr_c), _p(cc), W, H, float(nearest), 1.3, float(bg), S, cap, _p(rec), _p(depth),
                    _p(radius), _p(idx_sorted), _p(tile_range), _p(final_T), _p(ncontrib), _p(out),
                    host_t.data_ptr() + 4 * slot, _p(ws), ws.numel(), ev_arr, stream)
            if same_dev:
                _lib.check(lib.pxb_render_forward(*args), "pxb_render_forward")
            else:
                with torch.cuda.device(dev):
                    _lib.check(lib.pxb_render_forward(*args), "pxb_render_forward")
            _lib.count_launches("pxb_render_forward", W, H)
            if sync_free:
                _PENDING[dev.index].append((slot, cap, ckey))
                break
            n = _wait_count(word)
            ops.LAST_N[(dev.index, True)] = n
            if evs is not None:
                for k, name in enumerate(_FWD_STAGES):
                    if timer.wants(name):
                        timer.events.append((name, evs[k], evs[k + 1]))
            if n <= cap:
                if 2 * n < cap:  # shrink slowly when the scene got much lighter
                    _CAPACITY[ckey] = max(int(n * 1.25) + 65536, int(cap * 0.9))
                else:
                    _CAPACITY[ckey] = cap
                break
            # more intersections than the capacity: the lists were truncated, run again with room
            cap = _CAPACITY[ckey] = int(n * 1.25) + 65536
        ctx.save_for_backward(pos, sc, rot, sh, intr_c, extr_c, cc, rec, depth, radius, idx_sorted, tile_range, final_T,
                              ncontrib, *((op, sh_r) if sh_r is not None else ()))
        ctx.raw = sh_r is not None
        ctx.meta = (int(sh_degree), W, H, float(bg), int(with_depth), n_extra, S, Cc,
                    intr.shape, extr.shape, cam_center.shape, opacity.shape, shs.shape,
                    None if shs_rest is None else shs_rest.shape)
        ctx.mark_non_differentiable(radius)
        return out, radius

    @staticmethod
    def backward(ctx, d_out, _d_radius):
        saved = ctx.saved_tensors
        (pos, sc, rot, sh, intr, extr, cc, rec, depth, radius, idx_sorted, tile_range, final_T, ncontrib) = saved[:14]
        op_raw, sh_r = (saved[14], saved[15]) if ctx.raw else (None, None)
        sh_degree, W, H, bg, with_depth, n_extra, S, Cc, s_intr, s_extr, s_cc, s_op, s_sh, s_shr = ctx.meta
        dev = pos.device
        P = pos.shape[0]
        g = _f32(d_out, "dL_drendered")
        grec = torch.empty((P, S), dtype=torch.float32, device=dev)  # zeroed by the call
        # every P-sized gradient lives in ONE flat buffer (the 16-byte-aligned blocks first), so a
        # data-parallel caller exchanges them with a single collective (parallel.allreduce_step); a sink
        # may instead supply the buffers itself (symmetric memory) and ask for the factored SH gradient
        plan = None
        if _GRAD_SINK is not None and hasattr(_GRAD_SINK, "plan") and not ctx.raw:
            plan = _GRAD_SINK.plan(P, dev, cc, sh_degree)
        d_sh_rest = None
        if plan is not None:
            d_sh, d_rot, d_pos, d_sc, d_op, d_ndc, d_rgb = plan
        else:
            flat = _GRAD_SINK.next_buffer(61 * P) if (_GRAD_SINK is not None and not ctx.raw) else None
            if flat is None or flat.device != dev:
                flat = torch.empty(61 * P + 4, dtype=torch.float32, device=dev)
            if ctx.raw:
                # [features_rest 45P | pad to 4 | rotation 4P | features 3P | position 3P | scaling 3P | opacity P | ndc 2P]
                o = (45 * P + 3) // 4 * 4
                d_sh_rest = flat[0:45 * P].view(s_shr)
                d_rot = flat[o:o + 4 * P].view(P, 4)
                d_sh = flat[o + 4 * P:o + 7 * P].view(s_sh)
                o += 7 * P
            else:
                d_sh = flat[0:48 * P].view(P, 16, 3)
                d_rot = flat[48 * P:52 * P].view(P, 4)
                o = 52 * P
            d_pos = flat[o:o + 3 * P].view(P, 3)
            d_sc = flat[o + 3 * P:o + 6 * P].view(P, 3)
            d_op = flat[o + 6 * P:o + 7 * P]
            d_ndc = flat[o + 7 * P:o + 9 * P].view(P, 2)
            d_rgb = None
        d_extra = torch.empty(P, n_extra, dtype=torch.float32, device=dev) if n_extra > 0 else None
        need_cam = any(ctx.needs_input_grad[7:10])
        d_cam = torch.empty(19, dtype=torch.float32, device=dev) if need_cam else None
        timer = _lib._timer
        with torch.cuda.device(dev):
            evs, ev_arr = _stage_events(timer, _BWD_STAGES)
            _lib.check(lib.pxb_render_backward(
                P, sh_degree, _p(pos), _p(sc), _p(rot), _p(op_raw), _p(sh), _p(sh_r), n_extra, with_depth, _p(intr), _p(extr),
                _p(cc), W, H, bg, S, _p(rec), _p(depth), _p(radius), _p(idx_sorted), _p(tile_range), _p(final_T),
                _p(ncontrib), _p(g), _p(grec), _p(d_pos), _p(d_sc), _p(d_rot), _p(d_op),
                _p(None if d_rgb is not None else d_sh), _p(d_sh_rest), _p(d_rgb), _p(d_extra), _p(d_ndc), _p(d_cam), ev_arr,
                _raw_stream(dev.index)), "pxb_render_backward")
        _lib.count_launches("pxb_render_backward", W, H)
        if evs is not None:
            for k, name in enumerate(_BWD_STAGES):
                if timer.wants(name):
                    timer.events.append((name, evs[k], evs[k + 1]))
        d_intr = d_cam[0:4].reshape(s_intr) if ctx.needs_input_grad[7] else None
        d_extr = d_cam[4:16].reshape(s_extr) if ctx.needs_input_grad[8] else None
        d_cc = d_cam[16:19].reshape(s_cc) if ctx.needs_input_grad[9] else None
        return (d_pos, d_op.reshape(s_op), d_sc, d_rot, d_sh, d_sh_rest, d_extra, d_intr, d_extr, d_cc, d_ndc,
                None, None, None, None, None, None)


def fused_render(position, opacity, scaling, rotation, shs, intr, extr, cam_center, ndc, sh_degree, W, H, bg,
                 with_depth=False, extra=None, nearest=0.2, features_rest=None):
    """(features[C,H,W], radii[P]) through the fused path; ``extr`` is the 3x4 [R|T].  With ``features_rest`` the
    inputs are the point cloud's RAW parameters (``shs`` = features[P,1,3]; see MsplatRender.render_iter_raw)."""
    return _FusedRender.apply(position, opacity, scaling, rotation, shs, features_rest, extra, intr, extr, cam_center,
                              ndc, sh_degree, W, H, bg, with_depth, nearest)


class _CameraExtrinsics(torch.autograd.Function):
    """(qrot[4], tvec[3]) -> (extrinsic[4,4], center[3]): one single-warp kernel each way (csrc/camera.cu)."""

    @staticmethod
    def forward(ctx, qrot, tvec):
        q, t = _f32(qrot, "qrot").reshape(4), _f32(tvec, "tvec").reshape(3)
        E = torch.empty(4, 4, dtype=torch.float32, device=q.device)
        c = torch.empty(3, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            _lib.launch("pxb_camera_forward", _p(q), _p(t), _p(E), _p(c), _stream(q.device))
        ctx.save_for_backward(q, t)
        ctx.shapes = (qrot.shape, tvec.shape)
        ctx.set_materialize_grads(False)
        return E, c

    @staticmethod
    def backward(ctx, dE, dc):
        q, t = ctx.saved_tensors
        dE = None if dE is None else _f32(dE, "d_extrinsic")
        dc = None if dc is None else _f32(dc, "d_center")
        dq, dt = torch.empty_like(q), torch.empty_like(t)
        with torch.cuda.device(q.device):
            _lib.launch("pxb_camera_backward", _p(q), _p(t), _p(dE), _p(dc), _p(dq), _p(dt), _stream(q.device))
        return dq.reshape(ctx.shapes[0]), dt.reshape(ctx.shapes[1])


def camera_extrinsics(qrot: Tensor, tvec: Tensor):
    """``CameraModel.extrinsic_matrices`` / ``camera_centers`` for one view
    (pointrix/model/camera/camera_model.py:92-175): ``qrot[4]`` (w first; normalised inside, as the reference's
    ``F.normalize`` does), ``tvec[3]`` -> ``(extrinsic_matrix[4,4], camera_center[3])``, differentiable."""
    return _CameraExtrinsics.apply(qrot, tvec)


@register_renderer
class MsplatRender(BaseObject):
    """Render Gaussian point clouds with the B200-native msplat path."""

    @dataclass
    class Config:
        update_sh_iter: int = 1000
        max_sh_degree: int = 3
        render_depth: bool = False
        sync_free: bool = False  # not in the reference: no host wait in no_grad forwards (see check_sync_free)

    cfg: Config

    def setup(self, white_bg, device, **kwargs):
        self.sh_degree = 0
        self.device = device
        super().setup(white_bg, device, **kwargs)
        self.bg_color = 1.0 if white_bg else 0.0

    # ------------------------------------------------------------------
    def render_iter(self, height, width, extrinsic_matrix, intrinsic_params, camera_center, position, opacity,
                    scaling, rotation, shs, normals=None, extra_features: Optional[Dict[str, Tensor]] = None,
                    **kwargs) -> dict:
        """One view.  Extra per-Gaussian feature tensors ``[P,c]`` are blended as additional channels and
        returned under their name: ``normals`` (the keyword of examples/supervise/renderer.py:25-102) and
        any entry of the explicit ``extra_features`` dict (e.g. a 2-channel flow), in that order.  Every
        other keyword is ignored, as in the reference (msplat.py:51-64 swallows ``**kwargs``)."""
        if not position.is_cuda:
            raise RuntimeError("position must be a CUDA tensor")
        P = position.shape[0]
        extras: Dict[str, Tensor] = {}
        if normals is not None:
            extras["normals"] = normals
        if extra_features:
            extras.update(extra_features)
        for k, v in extras.items():
            if not (isinstance(v, Tensor) and v.dim() == 2 and v.shape[0] == P and v.is_floating_point()):
                raise ValueError(f"extra feature {k!r} must be a floating [P,c] tensor")
        extra = torch.cat(list(extras.values()), dim=-1) if extras else None
        extr = extrinsic_matrix[:3, :]
        intr = intrinsic_params.reshape(-1)[:4] if intrinsic_params.numel() != 4 else intrinsic_params
        n_ch = 3 + int(self.cfg.render_depth) + (0 if extra is None else extra.shape[1])
        ndc = torch.zeros(P, 2, dtype=torch.float32, device=position.device, requires_grad=True)
        try:
            ndc.retain_grad()
        except Exception:
            raise ValueError("ndc does not have grad")

        fused_ok = (shs.dim() == 3 and shs.shape[1] == 16 and shs.shape[2] == 3 and self.sh_degree <= 3
                    and n_ch <= ops.MAX_CH and P > 0)
        if fused_ok:
            global _SYNC_FREE
            _SYNC_FREE = bool(getattr(self.cfg, "sync_free", False))
            try:
                feats, radius = fused_render(position, opacity, scaling, rotation, shs, intr, extr, camera_center, ndc,
                                             self.sh_degree, width, height, self.bg_color, self.cfg.render_depth, extra)
            finally:
                _SYNC_FREE = False
        else:
            feats, radius = self._render_iter_ops(height, width, extr, intr, camera_center, position, opacity, scaling,
                                                  rotation, shs, extra, ndc)
        split, s = {}, 0
        names = [("rgb", 3)] + ([("depth", 1)] if self.cfg.render_depth else []) + [(k, v.shape[1]) for k, v in extras.items()]
        if len(names) == 1:  # rgb only: hand out the buffer itself (a slice costs a zeros+copy in autograd)
            split["rgb"] = feats
        else:
            for k, c in names:
                split[k] = feats[s:s + c]
                s += c
        return {"rendered_features_split": split, "uv_points": ndc, "visibility": radius > 0, "radii": radius}

    def render_iter_raw(self, height, width, extrinsic_matrix, intrinsic_params, camera_center, position, opacity,
                        scaling, rotation, features, features_rest, **kwargs) -> dict:
        """``render_iter`` on the point cloud's RAW parameters: ``opacity`` = logits,
        ``scaling`` = log-scales, ``rotation`` = un-normalised quaternions, ``features[P,1,3]`` /
        ``features_rest[P,15,3]`` -- what ``GaussianPointCloud`` stores, instead of what its ``get_opacity /
        get_scaling / get_rotation / get_shs`` properties materialise every iteration
        (pointrix/model/point_cloud/gaussian_points.py:70-86, base_model.py:79-85).  The activations run inside
        the fused kernels and the gradients arrive on the raw tensors.  Same returned dict as ``render_iter``."""
        if not position.is_cuda:
            raise RuntimeError("position must be a CUDA tensor")
        P = position.shape[0]
        if features.shape[0] != P or features.numel() != 3 * P or features_rest.numel() != 45 * P or self.sh_degree > 3:
            raise ValueError("render_iter_raw needs features[P,1,3], features_rest[P,15,3] and sh_degree <= 3")
        extr = extrinsic_matrix[:3, :]
        intr = intrinsic_params.reshape(-1)[:4] if intrinsic_params.numel() != 4 else intrinsic_params
        ndc = torch.zeros(P, 2, dtype=torch.float32, device=position.device, requires_grad=True)
        ndc.retain_grad()
        feats, radius = fused_render(position, opacity, scaling, rotation, features, intr, extr, camera_center, ndc,
                                     self.sh_degree, width, height, self.bg_color, self.cfg.render_depth, None,
                                     features_rest=features_rest)
        split = {"rgb": feats[0:3], "depth": feats[3:4]} if self.cfg.render_depth else {"rgb": feats}
        return {"rendered_features_split": split, "uv_points": ndc, "visibility": radius > 0, "radii": radius}

    def _render_iter_ops(self, height, width, extr, intr, camera_center, position, opacity, scaling, rotation, shs,
                         extra, ndc):
        """Operator-by-operator composition, the literal sequence of
        pointrix/model/renderer/msplat.py:94-151 (used for SH layouts / channel
        counts the fused kernels do not cover)."""
        direction = position - camera_center.reshape(1, 3)
        direction = direction / direction.norm(dim=1, keepdim=True)
        sh_coeff = shs.permute(0, 2, 1)
        sh_mask = torch.zeros_like(sh_coeff)
        sh_mask[..., :(self.sh_degree + 1) ** 2] = 1.0
        rgb = ops.compute_sh(sh_coeff * sh_mask, direction)
        rgb = (rgb + 0.5).clamp(min=0.0)
        uv, depth = ops.project_point(position, intr, extr, width, height, nearest=0.2)
        visible = depth != 0
        cov3d = ops.compute_cov3d(scaling, rotation, visible)
        conic, radius, tiles = ops.ewa_project(position, cov3d, intr, extr, uv, width, height, visible)
        ids, tile_range = ops.sort_gaussian(uv, depth, width, height, radius, tiles)
        cols = [rgb] + ([depth] if self.cfg.render_depth else []) + ([extra] if extra is not None else [])
        feats = ops.alpha_blending(uv, conic, opacity, torch.cat(cols, dim=-1), ids, tile_range, self.bg_color, width,
                                   height, ndc)
        return feats, radius

    # ------------------------------------------------------------------
    def render_batch(self, render_dict: dict) -> dict:
        """Loop the views of a batch (pointrix/model/renderer/msplat.py:160-213): only
        ``extrinsic_matrix`` and ``camera_center`` are sliced per view."""
        rendered_features: Dict[str, list] = {}
        uv_points, visibilitys, radii = [], [], []
        batched = ("extrinsic_matrix", "camera_center")
        for i in range(render_dict["extrinsic_matrix"].shape[0]):
            it = {k: (v[i, ...] if k in batched else v) for k, v in render_dict.items()}
            res = self.render_iter(**it)
            for name, feat in res["rendered_features_split"].items():
                rendered_features.setdefault(name, []).append(feat)
            uv_points.append(res["uv_points"])
            visibilitys.append(res["visibility"].unsqueeze(0))
            radii.append(res["radii"].unsqueeze(0))
        stacked = {k: torch.stack(v, dim=0) for k, v in rendered_features.items()}
        return {**stacked, "uv_points": uv_points, "visibility": torch.cat(visibilitys).any(dim=0),
                "radii": torch.cat(radii, 0).max(dim=0).values}

    def update_sh_degree(self, step):
        if step % self.cfg.update_sh_iter == 0:
            if self.sh_degree < self.cfg.max_sh_degree:
                self.sh_degree += 1

    def load_state_dict(self, state_dict):
        self.sh_degree = state_dict["sh_degree"]

    def state_dict(self):
        return {"sh_degree": self.sh_degree}
