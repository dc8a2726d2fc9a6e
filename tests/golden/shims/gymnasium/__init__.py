"""Stand-in for `gymnasium` (absent in this image) — TEST INFRASTRUCTURE ONLY.

Encodes the documented gymnasium>=0.26 behaviour the reference relies on:
`Env.np_random` is created lazily, and `Env.reset(seed=s)` REPLACES it with
`Generator(PCG64(SeedSequence(s)))` when a seed is given (base.py:142-143, 269).
"""
import numpy as np
from . import spaces  # noqa: F401
from .core import Env, Wrapper, ObservationWrapper  # noqa: F401
from .envs.registration import register, make  # noqa: F401
