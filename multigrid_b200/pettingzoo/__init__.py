"""Batched counterpart of `multigrid.pettingzoo` (reference multigrid/pettingzoo/__init__.py:38-115):
the PettingZoo `ParallelEnv` surface over a `BatchedMultiGridEnv`.

`reset` / `step` / `close` are the env's own; `possible_agents`, `observation_space(s)`,
`action_space(s)` as in the reference. `agents` -- in the reference the list of live agent ids,
empty once `is_done()` -- cannot be one list for `num_envs` envs: here it is the list of agent ids
that are live in AT LEAST ONE env of the batch (empty when every env is done), and `agent_mask`
gives the exact (num_envs, n) bool tensor `~terminated & ~is_done` the reference's property encodes.
Both synchronise with the GPU (they are host-side queries, not part of the step path).

pettingzoo is not installed in the build image: the wrapper subclasses `ParallelEnv` when it is
importable, and is a plain class with the same members otherwise.
"""
from __future__ import annotations

from typing import Any

import torch

from ..envs import make

try:  # pragma: no cover - pettingzoo is absent in the build image
    from pettingzoo import ParallelEnv
except Exception:  # noqa: BLE001
    ParallelEnv = object


class PettingZooWrapper(ParallelEnv):
    """pettingzoo/__init__.py:38-79 over a batched env."""

    def __init__(self, env):
        self.env = env
        self.reset = self.env.reset
        self.step = self.env.step
        self.close = self.env.close
        self.metadata = {}

    @property
    def agent_mask(self) -> torch.Tensor:
        """(num_envs, n) bool: agent i of env e is live (pettingzoo/__init__.py:52-56 per env)."""
        base = self.env.unwrapped
        terminated = torch.stack([agent.terminated for agent in base.agents], dim=1)
        return ~terminated & ~base.is_done()[:, None]

    @property
    def agents(self) -> list:
        live = self.agent_mask.any(dim=0).tolist()
        return [agent.index for agent, alive in zip(self.env.unwrapped.agents, live) if alive]

    @property
    def possible_agents(self) -> list:
        return [agent.index for agent in self.env.unwrapped.agents]

    @property
    def observation_spaces(self) -> dict:
        return {agent.index: agent.observation_space for agent in self.env.unwrapped.agents}

    @property
    def action_spaces(self) -> dict:
        return {agent.index: agent.action_space for agent in self.env.unwrapped.agents}

    @property
    def render_mode(self):
        return None  # rendering is out of scope of the batched engine

    def observation_space(self, agent_id):
        return self.env.unwrapped.agents[agent_id].observation_space

    def action_space(self, agent_id):
        return self.env.unwrapped.agents[agent_id].action_space


def to_pettingzoo_env(env_id: str, *wrappers, metadata: dict[str, Any] = {}):
    """pettingzoo/__init__.py:83-115, by registry id (see rllib.to_rllib_env)."""
    class PettingZooEnv(PettingZooWrapper):
        def __init__(self, *args, **kwargs):
            env = make(env_id, *args, **kwargs)
            for wrapper in wrappers:
                env = wrapper(env)
            super().__init__(env)

    PettingZooEnv.__name__ = f"PettingZoo_{env_id}"
    PettingZooEnv.metadata = metadata
    return PettingZooEnv
