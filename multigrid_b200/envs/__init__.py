"""Registry of the batched environments: the reference's ids (multigrid/envs/__init__.py:38-52)
over this package's layout generators.

    from multigrid_b200.envs import make
    env = make('MultiGrid-Empty-8x8-v0', agents=4, num_envs=65536)

`make(id, **kwargs)` splits kwargs the way the reference's env classes do: layout arguments
(`size`, `agent_start_pos`, `room_size`, ...) go to the layout, everything else
(`max_steps`, `agent_view_size`, `see_through_walls`, `allow_agent_overlap`, `joint_reward`,
`success_termination_mode`, `failure_termination_mode`) to `BatchedMultiGridEnv`, plus the
batch arguments `num_envs`, `device`, `auto_reset`, `pool_size`, `layout_seed`.
When gymnasium is importable the ids are also registered there (entry point = `make`).
"""
from __future__ import annotations

from ..env import BatchedMultiGridEnv
from ..layouts import (BlockedUnlockPickupLayout, EmptyLayout, LockedHallwayLayout, PlaygroundLayout,
                       RedBlueDoorsLayout)

# envs/redbluedoors.py:102-109
_RBD_DEFAULTS = dict(joint_reward=True, success_termination_mode='any', failure_termination_mode='any')

# id -> (layout class, layout kwargs, env defaults of the reference env class)
CONFIGURATIONS = {
    'MultiGrid-BlockedUnlockPickup-v0': (BlockedUnlockPickupLayout, {},
                                         dict(joint_reward=True, success_termination_mode='any')),
    'MultiGrid-Empty-5x5-v0': (EmptyLayout, {'size': 5}, {}),
    'MultiGrid-Empty-Random-5x5-v0': (EmptyLayout, {'size': 5, 'agent_start_pos': None}, {}),
    'MultiGrid-Empty-6x6-v0': (EmptyLayout, {'size': 6}, {}),
    'MultiGrid-Empty-Random-6x6-v0': (EmptyLayout, {'size': 6, 'agent_start_pos': None}, {}),
    'MultiGrid-Empty-8x8-v0': (EmptyLayout, {}, {}),
    'MultiGrid-Empty-16x16-v0': (EmptyLayout, {'size': 16}, {}),
    'MultiGrid-LockedHallway-2Rooms-v0': (LockedHallwayLayout, {'num_rooms': 2}, dict(joint_reward=True)),
    'MultiGrid-LockedHallway-4Rooms-v0': (LockedHallwayLayout, {'num_rooms': 4}, dict(joint_reward=True)),
    'MultiGrid-LockedHallway-6Rooms-v0': (LockedHallwayLayout, {'num_rooms': 6}, dict(joint_reward=True)),
    'MultiGrid-Playground-v0': (PlaygroundLayout, {}, {}),
    'MultiGrid-RedBlueDoors-6x6-v0': (RedBlueDoorsLayout, {'size': 6}, _RBD_DEFAULTS),
    'MultiGrid-RedBlueDoors-8x8-v0': (RedBlueDoorsLayout, {'size': 8}, _RBD_DEFAULTS),
}

# Reference ids that are not built yet: none (all 13 of multigrid/envs/__init__.py:38-52 are).
NOT_YET = ()

_LAYOUT_KEYS = {
    EmptyLayout: ('size', 'agent_start_pos', 'agent_start_dir'),
    BlockedUnlockPickupLayout: ('room_size',),
    PlaygroundLayout: ('room_size', 'num_rows', 'num_cols'),
    LockedHallwayLayout: ('num_rooms', 'room_size', 'max_hallway_keys', 'max_keys_per_room'),
    RedBlueDoorsLayout: ('size',),
}


def make(env_id: str, agents: int = 1, **kwargs) -> BatchedMultiGridEnv:
    if env_id in NOT_YET:
        raise NotImplementedError(f"{env_id}: its step() post-hook is not built yet")
    if env_id not in CONFIGURATIONS:
        raise KeyError(f"unknown environment id {env_id!r}")
    if not isinstance(agents, int):
        # gym.make(id, agents=Iterable[Agent]) (base.py:85-103, 156-180): the batched engine keeps no per-agent
        # objects, so an iterable only says how many agents there are; their settings must be the env's
        # (one view size and see_through_walls for all: gen_obs uses agents[0]'s anyway, base.py:364-365)
        try:
            agent_list = list(agents)
        except TypeError as exc:
            raise TypeError("agents must be an int or an iterable of agents") from exc
        for a in agent_list:
            vs = getattr(a, "view_size", None)
            if vs is not None and vs != kwargs.get("agent_view_size", 7):
                raise ValueError("per-agent view sizes are not supported: pass agent_view_size=<int> for all agents")
        agents = len(agent_list)
        if agents < 1:
            raise ValueError("at least one agent is needed")
    layout_cls, layout_kw, env_defaults = CONFIGURATIONS[env_id]
    layout_kw = dict(layout_kw)
    for key in _LAYOUT_KEYS[layout_cls]:
        if key in kwargs:
            layout_kw[key] = kwargs.pop(key)
    if 'max_steps' in kwargs and kwargs['max_steps'] is not None:
        layout_kw['max_steps'] = kwargs.pop('max_steps')
    layout = layout_cls(agents, **layout_kw)
    return BatchedMultiGridEnv(layout, **{**env_defaults, **kwargs})


try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium.envs.registration import register
    for _name in CONFIGURATIONS:
        register(id=_name, entry_point=make, kwargs={'env_id': _name})
except Exception:  # noqa: BLE001
    pass
