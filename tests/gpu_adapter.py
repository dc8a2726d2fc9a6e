"""Adapter giving the CUDA engine (through multigrid_b200.engine / the C ABI) the same
interface as oracle.OracleBatch, so the parity tests read the same for every implementation."""
from __future__ import annotations

import os

import numpy as np
import torch

from multigrid_b200 import _cabi
from multigrid_b200.engine import EngineConfig, StepEngine


def engine_config(cfg) -> EngineConfig:
    return EngineConfig(
        width=cfg.W, height=cfg.H, num_agents=cfg.n, view_size=cfg.V, max_steps=cfg.max_steps,
        see_through_walls=cfg.see_through_walls, allow_agent_overlap=cfg.allow_agent_overlap,
        joint_reward=cfg.joint_reward,
        success_termination_mode="any" if cfg.success_any else "all",
        failure_termination_mode="any" if cfg.failure_any else "all",
        hook=cfg.hook, hook_param=getattr(cfg, "hook_param", 0), auto_reset=cfg.auto_reset,
        layout_stride=cfg.layout_stride)


class GpuEngine:
    def __init__(self, cfg, grid, agents, pcg_state, pcg_inc, pool_grid=None, pool_agents=None,
                 layout_idx=None, step_count=None, host_path=False, fused=True, one_hot=None):
        self.cfg, self.host_path, self.fused = cfg, host_path, fused
        # MG_TEST_ONE_HOT=1 (tests/test_fused_one_hot.py): every fused device step ALSO asks the kernel for the
        # one-hot image (MgStepOut.one_hot) and checks it against the oracle's OneHotObsWrapper.one_hot
        self.one_hot = (os.environ.get("MG_TEST_ONE_HOT", "0") == "1") if one_hot is None else one_hot
        self.one_hot = self.one_hot and fused and not host_path
        grid = np.asarray(grid)
        self.B = grid.shape[0]
        if pool_grid is None:
            pool_grid, pool_agents = grid[:1], np.asarray(agents)[:1]
        self.eng = StepEngine(engine_config(cfg), self.B, "cuda:0", pool_grid, pool_agents)
        self.eng.load_state(grid, agents, step_count, pcg_state, pcg_inc, layout_idx)
        self.eng.obs_buf.fill_(0x55)
        if self.one_hot:
            self.eng.enable_one_hot()

    def _obs(self, buf):
        V = self.cfg.V
        buf = buf.cpu().numpy()
        assert (buf[:, :, 3 * V * V:] == 0).all(), "padding bytes must be zero"
        return buf[:, :, :3 * V * V].reshape(self.B, self.cfg.n, V, V, 3)

    def gen_obs(self):
        self.eng.gen_obs()
        return self._obs(self.eng.obs_buf)

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.int8)
        if self.host_path:
            h = self.eng.host_buffers()
            h["actions"].copy_(torch.from_numpy(actions))
            h = self.eng.step_host()
            self.eng.check_status()
            return (self._obs(h["obs"]), h["reward"].numpy().copy(),
                    h["terminated"].numpy().copy(), h["truncated"].numpy().copy())
        a = torch.from_numpy(actions).to("cuda:0")
        if self.fused:
            if self.one_hot:
                self.eng.one_hot.fill_(0x99)  # every byte must be written
            self.eng.step(a)
            if self.one_hot:
                from oracle.mg_oracle import one_hot
                np.testing.assert_array_equal(self.eng.one_hot.cpu().numpy(), one_hot(self._obs(self.eng.obs_buf)),
                                              err_msg="fused one-hot")
        else:  # mg_step + mg_gen_obs (only equivalent without post-hook / auto-reset)
            self.eng.step(a, fused=False)
            self.eng.gen_obs()
        self.eng.check_status()
        return (self._obs(self.eng.obs_buf), self.eng.reward.cpu().numpy(),
                self.eng.terminated.cpu().numpy(), self.eng.truncated.cpu().numpy())

    @property
    def grid(self):
        return self.eng.grid.cpu().numpy()

    @property
    def agents(self):
        return self.eng.agents.cpu().numpy()

    @property
    def step_count(self):
        return self.eng.step_count.cpu().numpy()

    @property
    def pcg_state(self):
        return self.eng.pcg_state.cpu().numpy().view(np.uint64)

    @property
    def layout_idx(self):
        return self.eng.layout_idx.cpu().numpy()
