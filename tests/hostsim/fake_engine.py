"""TEST INFRASTRUCTURE: a stand-in for multigrid_b200.engine.StepEngine that runs the kernels'
phase functions on the CPU (tests/hostsim) behind the same attribute/method surface, with CPU
torch tensors. Lets the host-side env layer (multigrid_b200/env.py) be tested without a GPU by
monkeypatching `multigrid_b200.env.StepEngine`. Never imported by the product."""
from __future__ import annotations

import numpy as np
import torch

from oracle import mg_oracle as O
from tests.hostsim.sim import SimEngine


class HostSimStepEngine:
    def __init__(self, cfg, num_envs, device="cpu", pool_grid=None, pool_agents=None):
        self.cfg, self.num_envs, self.device = cfg, int(num_envs), torch.device("cpu")
        E, n, W, H = self.num_envs, cfg.num_agents, cfg.width, cfg.height
        self._ocfg = O.OracleConfig(
            W=W, H=H, n=n, V=cfg.view_size, max_steps=cfg.max_steps,
            see_through_walls=cfg.see_through_walls, allow_agent_overlap=cfg.allow_agent_overlap,
            joint_reward=cfg.joint_reward, success_any=cfg.success_termination_mode == "any",
            failure_any=cfg.failure_termination_mode == "any", hook=cfg.hook, hook_param=cfg.hook_param,
            auto_reset=cfg.auto_reset, layout_stride=cfg.layout_stride)
        self._pending = dict(grid=np.zeros((E, W, H, 3), np.int8), agents=np.zeros((E, n, 8), np.int8),
                             pcg_state=np.zeros((E, 2), np.uint64), pcg_inc=np.zeros((E, 2), np.uint64),
                             layout_idx=np.zeros(E, np.int32), step_count=np.zeros(E, np.int32))
        self._pool = (pool_grid, pool_agents)
        self._sim = None

    def set_layout_pool(self, pool_grid, pool_agents):
        self._pool = (np.asarray(pool_grid, np.int8), np.asarray(pool_agents, np.int8))
        self._sim = None

    def gen_layout_pool_empty_random(self, rng_state, rng_inc, rng_buf=None):
        from tests.hostsim.sim import gen_layouts_empty_random
        cfg = self.cfg
        buf = np.zeros(len(rng_state), np.uint64) if rng_buf is None else rng_buf
        grid, agents, st, buf = gen_layouts_empty_random(cfg.width, cfg.height, cfg.num_agents, rng_state, rng_inc, buf)
        self.set_layout_pool(grid, agents)
        return np.array(st), np.array(buf)

    def gen_layout_pool_red_blue_doors(self, size, rng_state, rng_inc, rng_buf=None):
        from tests.hostsim.sim import gen_layouts_red_blue_doors
        buf = np.zeros(len(rng_state), np.uint64) if rng_buf is None else rng_buf
        grid, agents, st, buf = gen_layouts_red_blue_doors(size, self.cfg.num_agents, rng_state, rng_inc, buf)
        self.set_layout_pool(grid, agents)
        return np.array(st), np.array(buf)

    def gen_layout_pool_locked_hallway(self, num_rooms, room_size, mhk, mkpr, rng_state, rng_inc, rng_buf=None):
        from tests.hostsim.sim import gen_layouts_locked_hallway
        buf = np.zeros(len(rng_state), np.uint64) if rng_buf is None else rng_buf
        grid, agents, st, buf = gen_layouts_locked_hallway(num_rooms, room_size, mhk, mkpr, self.cfg.num_agents,
                                                           rng_state, rng_inc, buf)
        self.set_layout_pool(grid, agents)
        return np.array(st), np.array(buf)

    def gen_layout_pool_playground(self, room_size, num_rows, num_cols, rng_state, rng_inc, rng_buf, order_state,
                                   order_inc):
        from tests.hostsim.sim import gen_layouts_playground
        buf = np.zeros(len(rng_state), np.uint64) if rng_buf is None else rng_buf
        grid, agents, st, buf, ost = gen_layouts_playground(room_size, num_rows, num_cols, self.cfg.num_agents,
                                                            rng_state, rng_inc, buf, order_state, order_inc)
        self.set_layout_pool(grid, agents)
        return np.array(ost), np.zeros(len(rng_state), np.int32), np.array(st), np.array(buf)

    def gen_layout_pool_bup(self, room_size, rng_state, rng_inc, rng_buf, order_state, order_inc):
        from tests.hostsim.sim import gen_layouts_bup
        buf = np.zeros(len(rng_state), np.uint64) if rng_buf is None else rng_buf
        grid, agents, st, buf, ost, info = gen_layouts_bup(room_size, self.cfg.num_agents, rng_state, rng_inc, buf,
                                                           order_state, order_inc)
        self.set_layout_pool(grid, agents)
        return np.array(ost), np.array(info), np.array(st), np.array(buf)

    def load_state(self, grid=None, agents=None, step_count=None, pcg_state=None, pcg_inc=None,
                   layout_idx=None):
        self._snapshot()
        for k, v in dict(grid=grid, agents=agents, step_count=step_count, pcg_state=pcg_state,
                         pcg_inc=pcg_inc, layout_idx=layout_idx).items():
            if v is not None:
                self._pending[k] = np.asarray(v).astype(self._pending[k].dtype).reshape(self._pending[k].shape)
        self._sim = None

    def reset_from_pool(self, layout_idx=None):
        if layout_idx is not None:
            self.load_state(layout_idx=layout_idx)
        idx = self._pending["layout_idx"].astype(np.int64)
        self.load_state(grid=self._pool[0][idx], agents=self._pool[1][idx],
                        step_count=np.zeros(self.num_envs, np.int32))

    def _snapshot(self):
        if self._sim is not None:
            s = self._sim
            self._pending = dict(grid=s.grid.copy(), agents=s.agents.copy(), pcg_state=s.pcg_state.copy(),
                                 pcg_inc=s.pcg_inc.copy(), layout_idx=s.layout_idx.copy(),
                                 step_count=s.step_count.copy())

    def _engine(self):
        if self._sim is None:
            p = self._pending
            self._sim = SimEngine(self._ocfg, p["grid"], p["agents"], p["pcg_state"], p["pcg_inc"],
                                  pool_grid=self._pool[0], pool_agents=self._pool[1],
                                  layout_idx=p["layout_idx"], step_count=p["step_count"])
        return self._sim

    # tensors (zero-copy views of the simulator's numpy buffers)
    grid = property(lambda self: torch.from_numpy(self._engine().grid))
    agents = property(lambda self: torch.from_numpy(self._engine().agents))
    step_count = property(lambda self: torch.from_numpy(self._engine().step_count))
    layout_idx = property(lambda self: torch.from_numpy(self._engine().layout_idx))
    direction = property(lambda self: torch.from_numpy(self._engine().agents)[:, :, 0])

    def gen_obs(self):
        return torch.from_numpy(self._engine().gen_obs())

    def step(self, actions, chained=False):
        obs, rew, term, trunc = self._engine().step(actions.numpy())
        return (torch.from_numpy(obs), torch.from_numpy(rew), torch.from_numpy(term),
                torch.from_numpy(trunc))

    def check_status(self):
        pass
