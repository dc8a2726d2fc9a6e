"""Batched counterpart of `multigrid.rllib` (reference multigrid/rllib/__init__.py:44-111): the
RLlib `MultiAgentEnv` surface over a `BatchedMultiGridEnv`.

    from multigrid_b200.rllib import RLlibWrapper, to_rllib_env
    env = RLlibWrapper(make('MultiGrid-Empty-8x8-v0', agents=2, num_envs=4096))
    MyEnv = to_rllib_env('MultiGrid-Empty-8x8-v0', OneHotObsWrapper)
    env = MyEnv({'agents': 2, 'num_envs': 4096})

Same names and conventions as the reference (`agents`, `possible_agents`, `reset`, `step` adding
`'__all__'` to the termination / truncation dicts, `get_observation_space`, `get_action_space`);
every value keeps the leading `num_envs` axis of the batched engine, so `'__all__'` is an
(num_envs,) bool tensor (`all` over agents per env, computed on the device).

ray is not installed in the build image: when `ray.rllib` is importable the wrapper subclasses
`MultiAgentEnv` and the reference's ids are registered with `ray.tune` (each wrapped in
`OneHotObsWrapper`, as the reference does, rllib/__init__.py:110-111); otherwise it is a plain class
with the same methods.
"""
from __future__ import annotations

import functools

from ..envs import CONFIGURATIONS, make
from ..wrappers import OneHotObsWrapper

try:  # pragma: no cover - ray is absent in the build image
    from ray.rllib.env import MultiAgentEnv
    from ray.tune.registry import register_env
except Exception:  # noqa: BLE001
    MultiAgentEnv = object
    register_env = None


def _all(values: dict):
    """Per-env `all()` over the agents' tensors (rllib/__init__.py:61-62)."""
    return functools.reduce(lambda a, b: a & b, values.values())


class RLlibWrapper(MultiAgentEnv):
    """rllib/__init__.py:44-69 over a batched env."""

    def __init__(self, env):
        super().__init__()
        self.env = env
        self.agents = list(range(len(env.unwrapped.agents)))
        self.possible_agents = self.agents[:]

    def reset(self, *args, **kwargs):
        return self.env.reset(*args, **kwargs)

    def step(self, *args, **kwargs):
        obs, rewards, terminations, truncations, infos = self.env.step(*args, **kwargs)
        terminations['__all__'] = _all(terminations)
        truncations['__all__'] = _all(truncations)
        return obs, rewards, terminations, truncations, infos

    def get_observation_space(self, agent_index: int):
        return self.env.unwrapped.agents[agent_index].observation_space

    def get_action_space(self, agent_index: int):
        return self.env.unwrapped.agents[agent_index].action_space


def to_rllib_env(env_id: str, *wrappers, default_config: dict = {}):
    """rllib/__init__.py:72-105. The reference converts an env CLASS; the batched engine names its
    env families by registry id, so this takes the id (`make(env_id, **config)` builds the env)."""
    class RLlibEnv(RLlibWrapper):
        def __init__(self, config: dict = {}):
            config = {**default_config, **config}
            env = make(env_id, **config)
            for wrapper in wrappers:
                env = wrapper(env)
            super().__init__(env)

    RLlibEnv.__name__ = f"RLlib_{env_id}"
    return RLlibEnv


if register_env is not None:  # pragma: no cover
    for _name in CONFIGURATIONS:
        register_env(_name, to_rllib_env(_name, OneHotObsWrapper))
