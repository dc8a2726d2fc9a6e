"""Random dense "soup" states for differential tests (oracle vs kernels). Test infrastructure."""
from __future__ import annotations

import numpy as np

from oracle import mg_oracle as O


def random_layouts(cfg, K, rng):
    W, H, n = cfg.W, cfg.H, cfg.n
    grid = np.zeros((K, W, H, 3), np.int8)
    grid[..., 0] = O.EMPTY
    r = rng.random((K, W, H))
    kind = rng.integers(0, 10, (K, W, H))
    color = rng.integers(0, 6, (K, W, H))
    state = rng.integers(0, 3, (K, W, H))
    dense = r < 0.45

    def put(mask, t, c=None, s=None):
        grid[..., 0][mask] = t
        grid[..., 1][mask] = color[mask] if c is None else c
        grid[..., 2][mask] = 0 if s is None else s[mask]

    put(dense & (kind == 0), O.WALL, 5)
    put(dense & ((kind == 1) | (kind == 2)), O.DOOR, None, state)
    put(dense & ((kind == 3) | (kind == 4)), O.KEY)
    put(dense & (kind == 5), O.BALL)
    put(dense & (kind == 6), O.BOX)
    put(dense & (kind == 7), O.GOAL, 1)
    put(dense & (kind == 8), O.LAVA, 0)
    put(dense & (kind == 9), O.FLOOR)
    # outer wall ring (all registered envs have one)
    for sl in (np.s_[:, 0, :], np.s_[:, W - 1, :], np.s_[:, :, 0], np.s_[:, :, H - 1]):
        grid[sl] = (O.WALL, 5, 0)
    agents = np.zeros((K, n, 8), np.int8)
    agents[..., O.A_COLOR] = np.arange(n) % 6
    agents[..., O.A_DIR] = rng.integers(0, 4, (K, n))
    agents[..., O.A_X] = rng.integers(1, max(W - 1, 2), (K, n))
    agents[..., O.A_Y] = rng.integers(1, max(H - 1, 2), (K, n))
    agents[..., O.A_TERM] = rng.random((K, n)) < 0.1
    agents[..., O.A_CT] = O.EMPTY
    carrying = rng.random((K, n)) < 0.35
    ct = rng.integers(O.KEY, O.BOX + 1, (K, n))
    agents[..., O.A_CT][carrying] = ct[carrying]
    agents[..., O.A_CC][carrying] = rng.integers(0, 6, (K, n))[carrying]
    return grid, agents


def random_batch(cfg, B, seed, K=17):
    rng = np.random.default_rng(seed)
    grid, agents = random_layouts(cfg, B, rng)
    pool_grid, pool_agents = random_layouts(cfg, K, rng)
    pool_agents[..., O.A_TERM] = 0
    return dict(
        grid=grid, agents=agents,
        pcg_state=rng.integers(0, 2**63, (B, 2)).astype(np.uint64) * np.uint64(2) + np.uint64(1),
        pcg_inc=rng.integers(0, 2**63, (B, 2)).astype(np.uint64) * np.uint64(2) + np.uint64(1),
        pool_grid=pool_grid, pool_agents=pool_agents,
        layout_idx=rng.integers(0, K, B).astype(np.int32),
        step_count=rng.integers(0, 5, B).astype(np.int32),
    )
