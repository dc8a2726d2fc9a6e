"""Build the CUDA engine library in-tree: `python -m multigrid_b200.build [--force]`.

nvcc cross-compiles for sm_100a without a GPU. The .so is git-ignored but ships to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SRC_DIR = os.path.join(PKG_DIR, "csrc")
OUT = os.path.join(PKG_DIR, "_lib", "libmultigrid_b200.so")
SOURCES = ["mg_cabi.cu"]
DEPS = ["mg_cabi.cu", "mg_kernels.cuh", "mg_static.cuh", os.path.join(ROOT, "include", "multigrid_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    built = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(SRC_DIR, d)) > built for d in DEPS)


TRACE_OUT = os.path.join(PKG_DIR, "_lib", "libmultigrid_b200_trace.so")


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True builds the diagnostics variant (-DMG_TRACE, per-warp phase timestamps) next to the
    product library; tools/trace_timeline.py loads it through MG_LIB."""
    if trace:
        return _compile(TRACE_OUT, verbose, ["-DMG_TRACE"])
    if not force and not is_stale():
        return OUT
    return _compile(OUT, verbose, [])


def _compile(OUT: str, verbose: bool, extra: list) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
        "-I", os.path.join(ROOT, "include"), "-o", OUT,
    ] + extra + [os.path.join(SRC_DIR, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
