"""ctypes front-end for the C oracle (oracle/mg_oracle.c) -- TEST INFRASTRUCTURE, not product.

Same interface as `oracle.mg_oracle.OracleBatch`, but fast enough for 10^4..10^5 envs, and the
thing `bench.py` times for `cpu_baseline` / `--impl reference` (kind = "port").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .mg_oracle import OracleConfig

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmg_oracle.so")


class _Cfg(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "W", "H", "n", "V", "max_steps", "see_through_walls", "allow_overlap", "joint_reward",
        "success_any", "failure_any", "hook", "auto_reset", "layout_stride", "num_layouts",
        "obs_agent_stride", "hook_param")]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "mg_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()  # no-op when the library is newer than mg_oracle.c
        _lib = C.CDLL(LIB_PATH)
        _lib.mgo_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracle:
    def __init__(self, cfg: OracleConfig, grid, agents, pcg_state, pcg_inc, pool_grid=None,
                 pool_agents=None, layout_idx=None, step_count=None, nthreads: int = 1,
                 obs_agent_stride: int | None = None):
        self.cfg = cfg
        self.grid = np.ascontiguousarray(grid, dtype=np.int8).copy()
        self.agents = np.ascontiguousarray(agents, dtype=np.int8).copy()
        self.B = self.grid.shape[0]
        self.pcg_state = np.ascontiguousarray(pcg_state, dtype=np.uint64).copy()
        self.pcg_inc = np.ascontiguousarray(pcg_inc, dtype=np.uint64).copy()
        self.step_count = (np.zeros(self.B, np.int32) if step_count is None
                           else np.ascontiguousarray(step_count, dtype=np.int32).copy())
        self.layout_idx = (np.zeros(self.B, np.int32) if layout_idx is None
                           else np.ascontiguousarray(layout_idx, dtype=np.int32).copy())
        self.pool_grid = (self.grid[:1].copy() if pool_grid is None
                          else np.ascontiguousarray(pool_grid, dtype=np.int8))
        self.pool_agents = (self.agents[:1].copy() if pool_agents is None
                            else np.ascontiguousarray(pool_agents, dtype=np.int8))
        self.nthreads = nthreads
        self.stride = obs_agent_stride or 3 * cfg.V * cfg.V
        self.c = _Cfg(cfg.W, cfg.H, cfg.n, cfg.V, cfg.max_steps, int(cfg.see_through_walls),
                      int(cfg.allow_agent_overlap), int(cfg.joint_reward), int(cfg.success_any),
                      int(cfg.failure_any), int(cfg.hook), int(cfg.auto_reset),
                      int(cfg.layout_stride), int(self.pool_grid.shape[0]), self.stride,
                      int(cfg.hook_param))
        self.obs = np.zeros((self.B, cfg.n, self.stride), np.int8)
        self.reward = np.zeros((self.B, cfg.n), np.float64)
        self.terminated = np.zeros((self.B, cfg.n), np.uint8)
        self.truncated = np.zeros((self.B,), np.uint8)
        self.cell_flags = np.zeros((self.B, cfg.W * cfg.H), np.uint8)  # see mg_oracle.c handle_actions
        self.hook_state = np.zeros((self.B,), np.int32)

    def _obs_view(self):
        V = self.cfg.V
        return self.obs[:, :, :3 * V * V].reshape(self.B, self.cfg.n, V, V, 3)

    def gen_obs(self):
        rc = lib().mgo_gen_obs(C.byref(self.c), C.c_int64(self.B), _p(self.grid), _p(self.agents),
                               _p(self.obs), C.c_int(self.nthreads))
        assert rc == 0
        return self._obs_view()

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.int8)
        rc = lib().mgo_step_obs(
            C.byref(self.c), C.c_int64(self.B), _p(self.grid), _p(self.agents),
            _p(self.step_count), _p(self.pcg_state), _p(self.pcg_inc), _p(self.layout_idx),
            _p(self.pool_grid), _p(self.pool_agents), _p(actions), _p(self.obs), _p(self.reward),
            _p(self.terminated), _p(self.truncated), _p(self.cell_flags), _p(self.hook_state),
            C.c_int(self.nthreads))
        if rc == 1:
            raise ValueError("Unknown action")
        assert rc == 0
        return self._obs_view(), self.reward, self.terminated, self.truncated
