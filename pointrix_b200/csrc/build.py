"""Build libpointrix_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python pointrix_b200/csrc/build.py [--force] [--verbose]

The library has a plain C ABI (include/pointrix_b200.h) and links cudart
statically, so it depends on nothing but the CUDA driver.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libpointrix_b200.so")
SOURCES = ["preprocess.cu", "binning.cu", "blend.cu", "exchange.cu", "pipeline.cu", "loss.cu", "optim.cu", "camera.cu"]
HEADERS = ["common.cuh", "sh.cuh", os.path.join(ROOT, "include", "pointrix_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # the reference is built with --use_fast_math (msplat/setup.py:39): flush-to-zero is part of
    # its integer-exact arithmetic; the approximations themselves are issued explicitly (common.cuh)
    "--use_fast_math",
    "-Xcompiler", "-fPIC",
    "--expt-extended-lambda", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else (shutil.which("nvcc") or "nvcc")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, stats: bool = False) -> str:
    """stats=True builds the instrumented developer variant (-DPXB_STATS: pair-test counters in the
    blend kernels) next to the product library; the package never loads it unless PXB_LIBRARY says so."""
    lib = LIB.replace(".so", "_stats.so") if stats else LIB
    if not stats and not force and not needs_build():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build_stats" if stats else "build")
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *(["-DPXB_STATS"] if stats else []), "-c", os.path.join(HERE, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, stats="--stats" in sys.argv))
