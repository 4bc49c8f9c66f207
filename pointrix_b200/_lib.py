"""ctypes binding of libpointrix_b200.so (C ABI declared in include/pointrix_b200.h).

There is no CPU fallback: if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PXB_LIBRARY: developer override (e.g. the -DPXB_STATS build); never a fallback
LIB_PATH = os.environ.get("PXB_LIBRARY") or os.path.join(_HERE, "libpointrix_b200.so")

ERRORS = {-1: "bad argument", -2: "unsupported configuration", -3: "workspace too small", -4: "pointer not 16-byte aligned"}

p = C.c_void_p
i32 = C.c_int
i64 = C.c_longlong
f32 = C.c_float
f64 = C.c_double
sz = C.c_size_t


class AdamGroup(C.Structure):
    """struct pxb_adam_group (include/pointrix_b200.h)"""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("rows", C.c_longlong), ("width", C.c_int), ("param_stride", C.c_int), ("param_offset", C.c_int),
                ("grad_stride", C.c_int), ("grad_offset", C.c_int), ("step", C.c_int),
                ("lr", C.c_double)]


# name -> (restype, argtypes); mirrors include/pointrix_b200.h one to one
SIGNATURES = {
    "pxb_init": (i32, []),
    "pxb_project_point_forward": (i32, [i32, p, p, p, i32, i32, f32, f32, p, p, p]),
    "pxb_project_point_backward": (i32, [i32, p, p, p, p, p, p, p, p, p, p]),
    "pxb_compute_cov3d_forward": (i32, [i32, p, p, p, p, p]),
    "pxb_compute_cov3d_backward": (i32, [i32, p, p, p, p, p, p, p]),
    "pxb_ewa_project_forward": (i32, [i32, p, p, p, p, p, i32, i32, p, p, p, p, p]),
    "pxb_ewa_project_backward": (i32, [i32, p, p, p, p, p, p, p, p, p, p, p]),
    "pxb_compute_sh_forward": (i32, [i32, i32, i32, p, p, p, p, p]),
    "pxb_compute_sh_backward": (i32, [i32, i32, i32, p, p, p, p, p, p, p]),
    "pxb_bin_prepare_workspace_bytes": (sz, [i32]),
    "pxb_bin_sort_workspace_bytes": (sz, [i64, i32, i32]),
    "pxb_bin_prepare": (i32, [i32, p, p, p, p, p, sz, p]),
    "pxb_sort_gaussian": (i32, [i32, i64, p, p, i32, i32, p, p, p, i32, i32, p, p, p, p, sz, p, sz, p]),
    "pxb_record_stride": (i32, [i32]),
    "pxb_pack_records": (i32, [i32, p, p, p, p, i32, i32, i32, i32, p, p]),
    "pxb_unpack_grads": (i32, [i32, p, i32, i32, i32, i32, i32, p, p, p, p, p]),
    "pxb_blend_forward": (i32, [p, i32, i32, p, p, f32, i32, i32, p, p, p, p]),
    "pxb_blend_backward": (i32, [p, i32, i32, p, p, f32, i32, i32, p, p, p, p, p]),
    "pxb_blend_counters": (i32, [p]),
    "pxb_fused_forward": (i32, [i32, i32, p, p, p, p, p, p, p, i32, i32, p, p, p, i32, i32, f32, f32, i32, i32, p, p, p, p, p]),
    "pxb_render_workspace_bytes": (sz, [i32, i64, i32, i32]),
    "pxb_render_forward": (i32, [i32, i32, p, p, p, p, p, p, p, i32, i32, p, p, p, i32, i32, f32, f32, f32, i32, i64,
                                 p, p, p, p, p, p, p, p, p, p, sz, p, p]),
    "pxb_render_backward": (i32, [i32, i32, p, p, p, p, p, p, i32, i32, p, p, p, i32, i32, f32, i32, p, p, p, p, p, p, p,
                                  p, p, p, p, p, p, p, p, p, p, p, p, p, p]),
    "pxb_nvls_allreduce": (i32, [p, i64, i64, i32, i32, p]),
    "pxb_p2p_allreduce": (i32, [p, i64, i64, i32, i32, p]),
    "pxb_sh_grad_gather": (i32, [p, i64, i64, i32, i32, i32, p, p, p]),
    "pxb_fused_backward": (i32, [i32, i32, p, p, p, p, p, p, i32, i32, p, p, p, i32, i32, i32, p, p, p, p, p, p, p, p, p, p, p, p, p, p]),
    "pxb_loss_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "pxb_l1_ssim_forward": (i32, [i32, i32, i32, i32, p, p, p, p, p, p, sz, p]),
    "pxb_l1_ssim_loss_forward": (i32, [i32, i32, i32, i32, p, p, f32, p, p, p, sz, p]),
    "pxb_l1_ssim_backward": (i32, [i32, i32, i32, i32, p, p, p, p, p, i32, f32, f32, p, p]),
    "pxb_pixel_loss_forward": (i32, [i32, i32, i64, p, p, p, p, p, sz, p]),
    "pxb_pixel_loss_backward": (i32, [i32, i32, i64, p, p, p, p, p, p]),
    "pxb_camera_forward": (i32, [p, p, p, p, p]),
    "pxb_camera_backward": (i32, [p, p, p, p, p, p, p]),
    "pxb_adam_densify_step": (i32, [p, i32, f64, f64, f64, i32, p, p, f32, f32, p, p, p, p]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: pointrix_b200 has no CPU/PyTorch fallback. "
        "Build it with `python pointrix_b200/csrc/build.py` (needs nvcc)."
    )

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here means header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args

_initialised = False


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"pointrix_b200 {what}: {ERRORS.get(rc, rc)}")
    raise RuntimeError(f"pointrix_b200 {what}: CUDA error {rc}")


def ensure_init() -> None:
    global _initialised
    if not _initialised:
        check(lib.pxb_init(), "pxb_init")
        _initialised = True


# ---------------------------------------------------------------------------
# launch accounting: every C-ABI call of the package goes through launch().
# ---------------------------------------------------------------------------
# kernels launched per entry point (memsets not counted); sort adds its radix passes
KERNELS_PER_CALL = {
    "pxb_project_point_forward": 1, "pxb_project_point_backward": 1, "pxb_compute_cov3d_forward": 1,
    "pxb_compute_cov3d_backward": 1, "pxb_ewa_project_forward": 1, "pxb_ewa_project_backward": 1,
    "pxb_compute_sh_forward": 1, "pxb_compute_sh_backward": 1, "pxb_bin_prepare": 11, "pxb_sort_gaussian": 8,
    "pxb_pack_records": 1, "pxb_unpack_grads": 1, "pxb_blend_forward": 1, "pxb_blend_backward": 1,
    "pxb_fused_forward": 1, "pxb_fused_backward": 1,
    "pxb_adam_densify_step": 1, "pxb_l1_ssim_forward": 2, "pxb_l1_ssim_loss_forward": 2, "pxb_l1_ssim_backward": 1, "pxb_pixel_loss_forward": 2, "pxb_pixel_loss_backward": 1,
}


class KernelTimer:
    """CUDA-event timer around each C-ABI call (events are recorded on the stream the
    kernels are launched on: PyTorch's current stream)."""

    def __init__(self, stages=None):
        self.events = []  # (name, e0, e1)
        self.launches = 0
        # None: time every stage; a set of entry-point names: only those (each timed stage costs two
        # event records in the stream, ~5 us of step time)
        self.stages = None if stages is None else set(stages)

    def wants(self, name: str) -> bool:
        return self.stages is None or name in self.stages

    def summary(self):
        import collections

        tot, cnt = collections.OrderedDict(), collections.Counter()
        for name, e0, e1 in self.events:
            tot[name] = tot.get(name, 0.0) + e0.elapsed_time(e1)
            cnt[name] += 1
        return {k: {"ms_total": v, "calls": cnt[k], "ms_avg": v / cnt[k]} for k, v in tot.items()}


_timer = None
launch_count = 0


def set_timer(t):
    global _timer
    _timer = t


def _sort_kernels(W: int, H: int) -> int:
    nt = ((W + 15) // 16) * ((H + 15) // 16)
    return 2 + 3 * max(1, (max(nt - 1, 0).bit_length() + 7) // 8)


def count_launches(name: str, W: int, H: int) -> None:
    """Launch accounting of the whole-view entry points (called by the renderer)."""
    global launch_count
    # forward: fused per-Gaussian kernel, compaction + 4 x (scan, scatter), key emission + tile passes + ranges, blend
    n = (1 + 9 + _sort_kernels(W, H) + 1) if name == "pxb_render_forward" else 2
    launch_count += n
    if _timer is not None:
        _timer.launches += n


def launch(name: str, *args) -> None:
    global launch_count
    fn = getattr(lib, name)
    n = KERNELS_PER_CALL.get(name, 1)
    if name == "pxb_sort_gaussian":
        n = _sort_kernels(args[9], args[10]) if args[1] > 0 else 0
    launch_count += n
    if _timer is None or not _timer.wants(name):
        check(fn(*args), name)
        if _timer is not None:
            _timer.launches += n
        return
    import torch

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    _timer.events.append((name, e0, e1))
    _timer.launches += n
    check(rc, name)
