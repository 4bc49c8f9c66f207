"""Build recipe for ``oracle/_ref`` -- TEST INFRASTRUCTURE, never the product path.

Compiles the UNMODIFIED reference msplat CUDA extension from the sources where
they lie under ``/root/reference/msplat`` (no source is copied into this repo)
into ``oracle/_ref/msplat_ref_C*.so`` with the reference's own effective flags
(``-O3 --use_fast_math``, ``msplat/setup.py:39``; arch sm_100 as
``TORCH_CUDA_ARCH_LIST=10.0`` would give).  The reference's build system
(``setup.py``) is not run; the sources are handed to nvcc/g++ directly through
``torch.utils.cpp_extension.load`` (a ninja file it writes under
``oracle/_ref/build``).

The resulting module exposes the 12 entry points of
``msplat/msplat/src/ext.cpp:14-25`` and is used ONLY
  * by ``-m gpu`` parity tests as the bit-exact checker of integer outputs
    (radius, tiles, idx_sorted, tile_range, visibility),
  * by ``oracle/make_golden.py`` to generate committed golden vectors,
  * by ``bench.py`` to time the reference GPU path *beside* ours (``ref_gpu``).

``oracle/_ref/`` is git-ignored (built artefact) but NOT gpurun-ignored, so the
``.so`` travels to the GPU box; ``/root/reference`` does not exist there.
"""
from __future__ import annotations

import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("POINTRIX_REFERENCE", "/root/reference")
NAME = "msplat_ref_C"


def built_path() -> str | None:
    hits = sorted(glob.glob(os.path.join(OUT, NAME + "*.so")))
    return hits[0] if hits else None


def build(verbose: bool = False) -> str | None:
    """Compile the reference extension if its sources are present; return the .so path."""
    so = built_path()
    if so is not None:
        return so
    src_dir = os.path.join(REF, "msplat", "msplat", "src")
    if not os.path.isdir(src_dir):
        return None  # GPU box: only the prebuilt file is used
    from torch.utils.cpp_extension import load

    os.makedirs(os.path.join(OUT, "build"), exist_ok=True)
    sources = sorted(glob.glob(os.path.join(src_dir, "*.cu"))) + [os.path.join(src_dir, "ext.cpp")]
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    load(
        name=NAME,
        sources=sources,
        extra_include_paths=[
            os.path.join(REF, "msplat", "msplat", "include"),
            os.path.join(REF, "msplat", "third_party", "glm"),
        ],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math"],
        build_directory=os.path.join(OUT, "build"),
        is_python_module=False,
        verbose=verbose,
    )
    import shutil

    hits = glob.glob(os.path.join(OUT, "build", NAME + "*.so"))
    if not hits:
        raise RuntimeError("reference build produced no .so")
    dst = os.path.join(OUT, os.path.basename(hits[0]))
    shutil.copy2(hits[0], dst)
    return dst


def load_ref():
    """Import the prebuilt reference module (needs CUDA at call time, not at import)."""
    so = built_path()
    if so is None:
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", p)
