#!/usr/bin/env python
"""Stage the UNMODIFIED reference for timing on the GPU box. TEST / MEASUREMENT INFRASTRUCTURE.

    python oracle/make_ref.py            # run in the build container (needs /root/reference)

Copies the reference's package `/root/reference/multigrid` (pure Python + numba) to
`oracle/_ref/multigrid/`. `oracle/_ref/` is git-ignored (no reference source enters the history) but NOT
gpurun-ignored, so it travels to the GPU box, where `/root/reference` does not exist. The three third-party
imports the image lacks (gymnasium, aenum, pygame) come from the stand-ins in tests/golden/shims/ (SURVEY.md
Appendix B). `oracle/ref_runner.py` runs it; `bench.py --impl reference` and the `cpu_baseline` leg time it.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("MULTIGRID_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def stage(force: bool = False) -> str | None:
    """Returns the staged package directory, or None when the reference tree is absent (GPU box)."""
    src = os.path.join(REFERENCE, "multigrid")
    dst = os.path.join(DEST, "multigrid")
    if not os.path.isdir(src):
        return dst if os.path.isdir(dst) else None
    if os.path.isdir(dst) and not force:
        return dst
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Unmodified copy of /root/reference/multigrid staged by oracle/make_ref.py for CPU timing.\n"
                "Git-ignored; never edit, never import from the product package.\n")
    return dst


if __name__ == "__main__":
    out = stage(force="--force" in sys.argv)
    print(out if out else "reference tree not found: nothing staged")
