"""Batched counterparts of the reference's observation / agent wrappers (multigrid/wrappers.py).

They wrap a `BatchedMultiGridEnv` and keep its batched, on-device conventions:
  * `OneHotObsWrapper`  (wrappers.py:101-190): image -> uint8 (E, V, V, 21) one-hot, written by the fused step
    kernel itself (MgStepOut.one_hot) when it wraps the base env directly, else by the CUDA kernel `mg_one_hot`
    (no torch fallback);
  * `ImgObsWrapper`     (wrappers.py:61-98):  observations are the bare image tensors;
  * `SingleAgentWrapper`(wrappers.py:193-233): agent 0's items instead of per-agent dicts.
  * `FullyObsWrapper`   (wrappers.py:17-58):  the whole grid with all agents drawn in, the same
    (E, W, H, 3) int8 tensor for every agent, by the CUDA kernel `mg_full_obs`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi, spaces

ONE_HOT_CHANNELS = 21  # len(Type) + len(Color) + max(len(State), len(Direction)), wrappers.py:140-141


class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):  # gym.Wrapper forwards unknown attributes to the wrapped env
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def observation(self, obs):
        return obs

    def reset(self, *args, **kwargs):
        obs, infos = self.env.reset(*args, **kwargs)
        return self.observation(obs), infos

    def step(self, actions, **kwargs):
        obs, rewards, terminations, truncations, infos = self.env.step(actions, **kwargs)
        return self.observation(obs), rewards, terminations, truncations, infos


class OneHotObsWrapper(_Wrapper):
    """Images become one-hot: 11 type + 6 colour + 4 state/direction channels per cell."""

    def __init__(self, env):
        super().__init__(env)
        base = env.unwrapped
        E, n, V = base.num_envs, base.num_agents, base.agent_view_size
        # wrapping the base env itself: its fused step kernel emits the one-hot images from now on (`fused=False`
        # or a wrapped env in between: a separate pass over whatever image the observation holds)
        # (measured, profiles/r02_summary.md: step + one-hot 82 -> 55 us on Empty-8x8 x 65536)
        self._fused = env is base
        self._out = base.engine.enable_one_hot() if self._fused else torch.zeros(
            (E, n, V, V, ONE_HOT_CHANNELS), dtype=torch.uint8, device=base.device)
        for agent in base.agents:
            agent.observation_space["image"] = spaces.Box(low=0, high=1, shape=(V, V, ONE_HOT_CHANNELS),
                                                          dtype=np.uint8)

    def observation(self, obs):
        base = self.env.unwrapped
        eng = base.engine
        stream = C.c_void_p(torch.cuda.current_stream(base.device).cuda_stream)
        img0 = obs[0]["image"]
        if img0.data_ptr() == eng.obs_buf.data_ptr() and img0.shape[1:] == self._out.shape[2:4] + (3,):
            if self._fused and eng.one_hot is self._out:  # already written by the launch that produced `obs`
                for i in obs:
                    obs[i]["image"] = self._out[:, i]
                return obs
            # the wrapped env's own partial views: the whole observation buffer in one launch
            with torch.cuda.device(base.device):
                _cabi.check(eng.lib.mg_one_hot(base.agent_view_size, base.num_envs * base.num_agents,
                                               eng.obs_stride, eng.obs_buf.data_ptr(), self._out.data_ptr(), stream),
                            "mg_one_hot")
            for i in obs:
                obs[i]["image"] = self._out[:, i]
            return obs
        # images of another wrapper (e.g. FullyObsWrapper's whole-grid image): encode what `obs` holds
        # (wrappers.py:176-177 encodes obs[agent]['image'] whatever produced it)
        cache = {}
        for i in obs:
            img = obs[i]["image"]
            key = (img.data_ptr(), tuple(img.shape))
            if key not in cache:
                src = img.contiguous()
                cells = int(np.prod(src.shape[1:-1]))
                out = torch.empty(tuple(src.shape[:-1]) + (ONE_HOT_CHANNELS,), dtype=torch.uint8, device=src.device)
                with torch.cuda.device(base.device):
                    _cabi.check(eng.lib.mg_one_hot_cells(cells, src.shape[0], 3 * cells, src.data_ptr(), out.data_ptr(),
                                                         stream), "mg_one_hot_cells")
                cache[key] = out
            obs[i]["image"] = cache[key]
        return obs


class FullyObsWrapper(_Wrapper):
    """Every agent observes the whole grid (Grid.encode + every agent's encoding on its cell)."""

    def __init__(self, env):
        super().__init__(env)
        base = env.unwrapped
        self._out = torch.zeros((base.num_envs, base.width, base.height, 3), dtype=torch.int8,
                                device=base.device)
        for agent in base.agents:  # the reference declares (height, width, 3); the array is (W, H, 3)
            agent.observation_space["image"] = spaces.Box(low=0, high=255, shape=(base.height, base.width, 3),
                                                          dtype=np.int64)

    def observation(self, obs):
        base = self.env.unwrapped
        eng = base.engine
        with torch.cuda.device(base.device):
            _cabi.check(eng.lib.mg_full_obs(base.width, base.height, base.num_agents, base.num_envs,
                                            eng.cells.data_ptr(), eng.agents.data_ptr(), self._out.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream(base.device).cuda_stream)),
                        "mg_full_obs")
        for i in obs:
            obs[i]["image"] = self._out
        return obs


class ImgObsWrapper(_Wrapper):
    def observation(self, obs):
        return {i: o["image"] for i, o in obs.items()}


class SingleAgentWrapper(_Wrapper):
    """Agent 0's view of a (possibly multi-agent) env: `step(action)` acts for agent 0 only."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = env.unwrapped.agents[0].observation_space
        self.action_space = env.unwrapped.agents[0].action_space

    def reset(self, *args, **kwargs):
        return tuple(item[0] for item in self.env.reset(*args, **kwargs))

    def step(self, action):
        return tuple(item[0] for item in self.env.step({0: action}))
