"""The C-ABI library loads and exports every symbol include/pointrix_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pointrix_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pxb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pointrix_b200 import _lib

    names = _declared()
    assert len(names) >= 19
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_table_matches_header():
    from pointrix_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()


def test_argument_counts_match_header():
    from pointrix_b200 import _lib

    src = open(os.path.join(ROOT, "include", "pointrix_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, params in re.findall(r"\b(pxb_\w+)\s*\(([^)]*)\)\s*;", src):
        n = 0 if params.strip() in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), name


def test_pure_host_entry_points():
    """Entry points that never touch the device can be called without a GPU."""
    from pointrix_b200 import _lib

    assert [_lib.lib.pxb_record_stride(c) for c in (1, 2, 3, 6, 7, 10, 18, 26, 27)] == [8, 8, 12, 12, 16, 16, 24, 32, -1]
    a = _lib.lib.pxb_bin_sort_workspace_bytes(10000, 1920, 1080)
    b = _lib.lib.pxb_bin_sort_workspace_bytes(20000, 1920, 1080)
    assert 0 < a < b
    assert 0 < _lib.lib.pxb_bin_prepare_workspace_bytes(1000) < _lib.lib.pxb_bin_prepare_workspace_bytes(100000)


def test_header_is_plain_c_and_links(tmp_path):
    """include/pointrix_b200.h compiles as C (not only C++) and a C program links against the library
    and calls its pure-host entry points -- the boundary really is a C ABI."""
    import shutil
    import subprocess

    from pointrix_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest

        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text(
        '#include <stdio.h>\n#include "pointrix_b200.h"\n'
        "int main(void) {\n"
        '  printf("%d %d %zu %zu\\n", pxb_record_stride(3), pxb_record_stride(27),\n'
        "         pxb_loss_workspace_bytes(1, 3, 1080, 1920), pxb_render_workspace_bytes(1000, 100000, 640, 480));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "t"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert out[0] == "12" and out[1] == "-1"
    assert int(out[2]) == 2 * 60 * 34 * 3 * 4  # two fp32 partial sums per 32x32 tile of each of the 3 planes
    assert int(out[3]) > 0


def test_argument_errors_are_return_codes():
    """Bad arguments come back as PXB_ERR_* before anything touches the device (no GPU needed)."""
    import ctypes as C

    from pointrix_b200 import _lib

    L = _lib.lib
    null, one = C.c_void_p(0), C.c_void_p(16)
    assert L.pxb_l1_ssim_forward(0, 3, 8, 8, one, one, null, one, one, one, 1 << 20, null) == -1       # B = 0
    assert L.pxb_l1_ssim_forward(1, 3, 8, 8, one, one, null, one, one, one, 4, null) == -3             # workspace too small
    assert L.pxb_l1_ssim_forward(70000, 1, 8, 8, one, one, null, one, one, one, 1 << 30, null) == -2   # B*C > 65535
    assert L.pxb_l1_ssim_loss_forward(1, 3, 8, 8, one, one, 0.2, null, null, one, 1 << 20, null) == -1  # no output
    assert L.pxb_l1_ssim_backward(1, 3, 8, 8, one, one, null, null, null, 0, 1.0, 1.0, one, null) == -1  # no maps
    assert L.pxb_pixel_loss_forward(3, 1, 10, one, one, null, one, one, 1 << 20, null) == -1            # unknown mode
    assert L.pxb_pixel_loss_forward(1, 1, 1 << 20, one, one, null, one, one, 16, null) == -3
    assert L.pxb_sh_grad_gather(null, 0, 0, 2, 10, 3, one, one, null) == -1                            # no peers
    arr = (C.c_void_p * 2)(16, 16)
    assert L.pxb_sh_grad_gather(arr, 0, 0, 17, 10, 3, one, one, null) == -1                            # world > 16
    assert L.pxb_sh_grad_gather(arr, 0, 0, 2, 10, 4, one, one, null) == -2                             # SH degree > 3
    assert L.pxb_sh_grad_gather(arr, 0, 0, 2, 10, 3, one, C.c_void_p(20), null) == -4                  # d_shs misaligned
    assert L.pxb_p2p_allreduce(arr, 6, 0, 0, 2, null) == -4                                            # n_f32 % (4*world)
    assert L.pxb_nvls_allreduce(null, 8, 0, 0, 2, null) == -1
    assert L.pxb_render_forward(0, 3, *([null] * 7), 0, 0, null, null, null, 8, 8, 0.2, 1.3, 1.0, 12, 100,
                                *([null] * 10), 0, null, null) == -1                                   # P = 0
    assert L.pxb_fused_backward(10, 5, *([one] * 4), null, null, 0, 0, one, one, one, 8, 8, 12, *([one] * 14), null) == -2  # SH degree
    assert L.pxb_fused_backward(10, 3, *([one] * 4), null, null, 0, 0, one, one, one, 8, 8, 12, one, one, one, one, one, one, one,
                                null, null, null, one, one, one, null) == -1                           # neither d_shs nor d_rgb
    assert L.pxb_fused_backward(10, 3, *([one] * 3), null, one, one, 0, 0, one, one, one, 8, 8, 12, one, one, one, one, one, one, one,
                                one, one, null, one, one, one, null) == -1                             # raw mode without the logits
    assert L.pxb_camera_forward(null, one, one, one, null) == -1
    assert [L.pxb_loss_workspace_bytes(0, 3, 8, 8), L.pxb_loss_workspace_bytes(1, 1, 1, 1)] == [0, 8]
    grp = (_lib.AdamGroup * 1)(_lib.AdamGroup(16, 16, 16, 16, 10, 3, 3, 0, 2, 0, 1, 1e-3))        # grad_stride < offset + width
    assert L.pxb_adam_densify_step(C.cast(grp, C.c_void_p), 1, 0.9, 0.999, 1e-15, 0, null, null, 1.0, 1.0, null, null, null, null) == -1
    grp = (_lib.AdamGroup * 1)(_lib.AdamGroup(16, 16, 16, 16, 10, 3, 3, 0, 3, 0, 0, 1e-3))        # step < 1
    assert L.pxb_adam_densify_step(C.cast(grp, C.c_void_p), 1, 0.9, 0.999, 1e-15, 0, null, null, 1.0, 1.0, null, null, null, null) == -1
    assert L.pxb_adam_densify_step(null, 9, 0.9, 0.999, 1e-15, 0, null, null, 1.0, 1.0, null, null, null, null) == -1    # > 8 groups
    assert L.pxb_adam_densify_step(null, 0, 0.9, 0.999, 1e-15, 5, one, null, 1.0, 1.0, one, one, one, null) == -1        # no radii
    assert L.pxb_adam_densify_step(null, 0, 0.9, 0.999, 1e-15, 0, null, null, 1.0, 1.0, null, null, null, null) == 0     # nothing to do
