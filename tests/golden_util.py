"""Helpers shared by the golden-fixture tests (loading the .npz files recorded from the reference)."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROLLOUT_CASES = sorted(
    os.path.splitext(os.path.basename(p))[0]
    for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
    if os.path.basename(p) not in ("obs_random.npz", "pcg64_kat.npz", "one_hot_kat.npz", "full_obs_kat.npz")
)


def load_case(name):
    d = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    meta = {k[5:]: d[k].item() for k in d.files if k.startswith("meta_")}
    return d, meta
