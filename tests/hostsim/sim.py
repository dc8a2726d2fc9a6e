"""TEST INFRASTRUCTURE: ctypes front-end of the CPU thread-by-thread run of the CUDA phase
functions (tests/hostsim/hostsim.cpp). Lets the kernel logic be diffed against the oracle in a
container without a GPU. Not importable from the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from multigrid_b200 import _cabi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libhostsim.so")
MODE_OBS, MODE_STEP, MODE_STEP_OBS = 0, 1, 2


def build():
    srcs = [os.path.join(HERE, "hostsim.cpp"),
            os.path.join(ROOT, "multigrid_b200", "csrc", "mg_kernels.cuh"),
            os.path.join(ROOT, "include", "multigrid_b200.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call([
        "/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=undefined",
        "-fno-sanitize-recover=all", "-I", os.path.join(ROOT, "include"), "-o", LIB, srcs[0]])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.sim_run.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def aligned(shape, dtype):
    """numpy array whose data pointer is 16-byte aligned (the ABI requires it)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    buf = np.zeros(nbytes + 16, dtype=np.uint8)
    off = (-buf.ctypes.data) % 16
    return buf[off:off + nbytes].view(dtype).reshape(shape)


def aligned_copy(a, dtype):
    out = aligned(np.shape(a), dtype)
    out[...] = a
    return out


def pack_cells(grid3):
    """(K,W,H,3) int8 -> (K,W+1,H+1) uint32 cell words with wall sentinels (numpy restatement of
    mg_pack_grid; bit 31 = opaque = wall or non-open door)."""
    g = np.asarray(grid3).astype(np.uint32)
    K, W, H, _ = g.shape
    t, c, s = g[..., 0], g[..., 1], g[..., 2]
    opaque = ((t == 2) | ((t == 4) & (s != 0))).astype(np.uint32)
    cells = np.full((K, W + 1, H + 1), 2 | (5 << 8) | (1 << 31), np.uint32)
    cells[:, :W, :H] = t | (c << 8) | (s << 16) | (opaque << 31)
    return cells


def unpack_cells(cells, W, H):
    w = np.asarray(cells)[:, :W, :H]
    return np.stack([w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff], -1).astype(np.int8)


def gen_layouts_empty_random(W, H, n, rng_state, rng_inc, rng_buf):
    """CPU run of the layout kernel's function. Returns (grid (K,W,H,3) int8, agents (K,n,8), state, buf)."""
    K = len(rng_state)
    st, inc = aligned_copy(rng_state, np.uint64), aligned_copy(rng_inc, np.uint64)
    buf = aligned_copy(rng_buf, np.uint64)
    cells, agents = aligned((K, W + 1, H + 1), np.uint32), aligned((K, n, 8), np.int8)
    rc = lib().sim_gen_layouts_empty_random(C.c_int(W), C.c_int(H), C.c_int(n), C.c_int64(K), _p(st), _p(inc),
                                            _p(buf), _p(cells), _p(agents))
    assert rc == 0, rc
    assert (cells[:, W, :] == cells[:, 0, :]).all() and (cells[:, :, H] == cells[:, :, 0]).all()  # sentinels = walls
    return unpack_cells(cells, W, H), agents, st, buf


def gen_layouts_red_blue_doors(size, n, rng_state, rng_inc, rng_buf):
    K, W, H = len(rng_state), 2 * size, size
    st, inc, buf = aligned_copy(rng_state, np.uint64), aligned_copy(rng_inc, np.uint64), aligned_copy(rng_buf, np.uint64)
    cells, agents = aligned((K, W + 1, H + 1), np.uint32), aligned((K, n, 8), np.int8)
    rc = lib().sim_gen_layouts_red_blue_doors(C.c_int(size), C.c_int(n), C.c_int64(K), _p(st), _p(inc), _p(buf),
                                              _p(cells), _p(agents))
    assert rc == 0, rc
    return unpack_cells(cells, W, H), agents, st, buf


def gen_layouts_locked_hallway(num_rooms, S, mhk, mkpr, n, rng_state, rng_inc, rng_buf):
    K, W, H = len(rng_state), 3 * (S - 1) + 1, (num_rooms // 2) * (S - 1) + 1
    st, inc, buf = aligned_copy(rng_state, np.uint64), aligned_copy(rng_inc, np.uint64), aligned_copy(rng_buf, np.uint64)
    cells, agents = aligned((K, W + 1, H + 1), np.uint32), aligned((K, n, 8), np.int8)
    rc = lib().sim_gen_layouts_locked_hallway(C.c_int(num_rooms), C.c_int(S), C.c_int(mhk), C.c_int(mkpr), C.c_int(n),
                                              C.c_int64(K), _p(st), _p(inc), _p(buf), _p(cells), _p(agents))
    assert rc == 0, rc
    return unpack_cells(cells, W, H), agents, st, buf


def gen_layouts_playground(S, rows, cols, n, rng_state, rng_inc, rng_buf, order_state, order_inc):
    K, W, H = len(rng_state), cols * (S - 1) + 1, rows * (S - 1) + 1
    st, inc, buf = aligned_copy(rng_state, np.uint64), aligned_copy(rng_inc, np.uint64), aligned_copy(rng_buf, np.uint64)
    ost, oinc = aligned_copy(order_state, np.uint64), aligned_copy(order_inc, np.uint64)
    cells, agents = aligned((K, W + 1, H + 1), np.uint32), aligned((K, n, 8), np.int8)
    obuf = aligned((K,), np.uint64)
    rc = lib().sim_gen_layouts_playground(C.c_int(S), C.c_int(rows), C.c_int(cols), C.c_int(n), C.c_int64(K), _p(st),
                                          _p(inc), _p(buf), _p(ost), _p(oinc), _p(obuf), _p(cells), _p(agents))
    assert rc == 0, rc
    gen_layouts_playground.order_buf = obuf  # (buffered half of the order streams, for sim_refresh_done_layouts)
    return unpack_cells(cells, W, H), agents, st, buf, ost


def gen_layouts_bup(S, n, rng_state, rng_inc, rng_buf, order_state, order_inc):
    """CPU run of the BUP layout function. Returns (grid, agents, rng_state, rng_buf, order_state, box colours)."""
    K, W, H = len(rng_state), 2 * (S - 1) + 1, S
    st, inc, buf = aligned_copy(rng_state, np.uint64), aligned_copy(rng_inc, np.uint64), aligned_copy(rng_buf, np.uint64)
    ost, oinc = aligned_copy(order_state, np.uint64), aligned_copy(order_inc, np.uint64)
    cells, agents, info = aligned((K, W + 1, H + 1), np.uint32), aligned((K, n, 8), np.int8), aligned((K,), np.int32)
    obuf = aligned((K,), np.uint64)
    rc = lib().sim_gen_layouts_bup(C.c_int(S), C.c_int(n), C.c_int64(K), _p(st), _p(inc), _p(buf), _p(ost), _p(oinc),
                                   _p(obuf), _p(cells), _p(agents), _p(info))
    assert rc == 0, rc
    gen_layouts_bup.order_buf = obuf  # (buffered half of the order streams, for sim_refresh_done_layouts)
    return unpack_cells(cells, W, H), agents, st, buf, ost, info


class SimEngine:
    """Mirror of oracle.OracleBatch's interface on top of the host-simulated kernels."""

    def __init__(self, cfg, grid, agents, pcg_state, pcg_inc, pool_grid=None, pool_agents=None,
                 layout_idx=None, step_count=None, forced_group=0, generic=0, split=False, static=False):
        self.cfg, self.forced_group, self.generic, self.split = cfg, forced_group, generic, split
        self.cells = aligned_copy(pack_cells(grid), np.uint32)
        self.agents = aligned_copy(agents, np.int8)
        self.B = self.cells.shape[0]
        self.pcg_state = aligned_copy(pcg_state, np.uint64)
        self.pcg_inc = aligned_copy(pcg_inc, np.uint64)
        self.step_count = aligned((self.B,), np.int32)
        if step_count is not None:
            self.step_count[...] = step_count
        self.layout_idx = aligned((self.B,), np.int32)
        if layout_idx is not None:
            self.layout_idx[...] = layout_idx
        self.pool_grid = aligned_copy(self.cells[:1] if pool_grid is None else pack_cells(pool_grid), np.uint32)
        self.pool_agents = aligned_copy(self.agents[:1] if pool_agents is None else pool_agents, np.int8)
        # static path: observation slots of the table's entry stride (what StepEngine picks for a static layout)
        self.stride = (_cabi.obs_agent_stride(cfg.V) + 15) & ~15 if static else _cabi.obs_agent_stride(cfg.V)
        flags = ((_cabi.FLAG_SEE_THROUGH_WALLS if cfg.see_through_walls else 0)
                 | (_cabi.FLAG_ALLOW_OVERLAP if cfg.allow_agent_overlap else 0)
                 | (_cabi.FLAG_JOINT_REWARD if cfg.joint_reward else 0)
                 | (_cabi.FLAG_SUCCESS_ANY if cfg.success_any else 0)
                 | (_cabi.FLAG_FAILURE_ANY if cfg.failure_any else 0)
                 | (_cabi.FLAG_AUTO_RESET if cfg.auto_reset else 0))
        self.c = _cabi.MgConfig(cfg.W, cfg.H, cfg.n, cfg.V, cfg.max_steps, flags, cfg.hook,
                                self.stride, self.pool_grid.shape[0], cfg.layout_stride,
                                getattr(cfg, "hook_param", 0))
        self.obs = aligned((self.B, cfg.n, self.stride), np.int8)
        self.obs[...] = 0x55
        self.reward = aligned((self.B, cfg.n), np.float64)
        self.terminated = aligned((self.B, cfg.n), np.uint8)
        self.truncated = aligned((self.B,), np.uint8)
        self.status = aligned((1,), np.int32)
        self.hook_state = aligned((self.B,), np.int32)
        self.state = _cabi.MgState(_p(self.cells).value, _p(self.agents).value,
                                   _p(self.step_count).value, _p(self.pcg_state).value,
                                   _p(self.pcg_inc).value, _p(self.layout_idx).value,
                                   _p(self.pool_grid).value, _p(self.pool_agents).value,
                                   _p(self.hook_state).value)
        # the fused one-hot image (MgStepOut.one_hot), poisoned so that every byte must be written
        self.one_hot = aligned((self.B, cfg.n, cfg.V, cfg.V, 21), np.uint8)
        self.one_hot[...] = 0x99
        self.out = _cabi.MgStepOut(_p(self.obs).value, _p(self.reward).value,
                                   _p(self.terminated).value, _p(self.truncated).value,
                                   _p(self.status).value, _p(self.one_hot).value)
        self.c_static = None
        if static:  # MG_FLAG_STATIC_GRID: the caller's promise must hold (checked here like StepEngine does)
            from multigrid_b200.engine import static_layout_ok
            pg3 = unpack_cells(self.pool_grid, cfg.W, cfg.H)
            assert cfg.hook == 0 and static_layout_ok(pg3, self.pool_agents), "not a static layout"
            assert (self.cells == self.pool_grid[0]).all(), "grids differ from the layout"
            assert static_layout_ok(pg3, self.agents.reshape(1, -1, 8)[:, :, :]), "agents break the promise"
            ts = (self.stride + 15) & ~15
            self.static_obs = aligned((cfg.W * cfg.H * 4 * (ts + 4),), np.uint8)
            self.static_obs[...] = 0x77
            rc = lib().sim_build_static_obs(C.byref(self.c), _p(self.pool_grid), _p(self.static_obs))
            assert rc == 0, rc
            self.state.static_obs = _p(self.static_obs).value
            self.c_static = _cabi.MgConfig(cfg.W, cfg.H, cfg.n, cfg.V, cfg.max_steps, flags | _cabi.FLAG_STATIC_GRID,
                                           cfg.hook, self.stride, self.pool_grid.shape[0], cfg.layout_stride,
                                           getattr(cfg, "hook_param", 0))

    @property
    def grid(self):
        return unpack_cells(self.cells, self.cfg.W, self.cfg.H)

    def _run(self, mode, actions=None, out=None, T=1, direction=None):
        c = self.c_static if (self.c_static is not None and mode == MODE_STEP_OBS and T == 1) else self.c
        rc = lib().sim_run(C.c_int(mode), C.byref(c), C.c_int64(self.B), C.byref(self.state),
                           _p(actions), C.byref(self.out if out is None else out), C.c_int(self.forced_group),
                           C.c_int(self.generic), C.c_int(T), _p(direction))
        assert rc == 0, rc

    def rollout(self, actions):
        """T steps in one run of the kernel's step loop (mg_rollout). actions (T,B,n)."""
        actions = aligned_copy(actions, np.int8)
        T, cfg, V = actions.shape[0], self.cfg, self.cfg.V
        obs = aligned((T, self.B, cfg.n, self.stride), np.int8)
        obs[...] = 0x55
        rew, term = aligned((T, self.B, cfg.n), np.float64), aligned((T, self.B, cfg.n), np.uint8)
        trunc, dirs = aligned((T, self.B), np.uint8), aligned((T, self.B, cfg.n), np.int8)
        out = _cabi.MgStepOut(_p(obs).value, _p(rew).value, _p(term).value, _p(trunc).value,
                              _p(self.status).value, None)
        self._run(MODE_STEP_OBS, actions, out=out, T=T, direction=dirs)
        if self.status[0] & 1:
            raise ValueError("Unknown action")
        assert (obs[..., 3 * V * V:] == 0).all(), "padding bytes must be zero"
        return obs[..., :3 * V * V].reshape(T, self.B, cfg.n, V, V, 3), dirs, rew, term, trunc

    def _obs_view(self):
        V = self.cfg.V
        assert (self.obs[:, :, 3 * V * V:] == 0).all(), "padding bytes must be zero"
        return self.obs[:, :, :3 * V * V].reshape(self.B, self.cfg.n, V, V, 3)

    def gen_obs(self):
        self._run(MODE_OBS)
        return self._obs_view()

    def step(self, actions):
        actions = aligned_copy(actions, np.int8)
        if self.split and not self.cfg.auto_reset:
            # mg_step then mg_gen_obs: with a post-hook the observation must see pre-hook
            # termination, which only the fused kernel provides; split is used hook-free.
            self._run(MODE_STEP, actions)
            self._run(MODE_OBS)
        else:
            self.one_hot[...] = 0x99
            self._run(MODE_STEP_OBS, actions)
            # every fused step of every hostsim test also checks the fused one-hot image (MgStepOut.one_hot)
            # against the oracle's restatement of OneHotObsWrapper.one_hot applied to the observations
            from oracle.mg_oracle import one_hot
            np.testing.assert_array_equal(self.one_hot, one_hot(self._obs_view()), err_msg="fused one-hot")
        if self.status[0] & 1:
            raise ValueError("Unknown action")
        assert not (self.status[0] & 4), "static-grid promise violated"
        return self._obs_view(), self.reward, self.terminated, self.truncated
