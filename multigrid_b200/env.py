"""`BatchedMultiGridEnv`: the reference's `MultiGridEnv` surface over `num_envs` lock-stepped envs.

Mirrors multigrid/base.py:36-841 for the hot path -- `reset(seed)`, `step(actions)` returning
per-agent dicts keyed `0..n-1`, `agents[i]`, `grid.state`, `agent_states`, `step_count`,
`max_steps`, `is_done()`, `observation_space` / `action_space` -- with one difference: every
value gains a leading `num_envs` axis and lives on the GPU (`image: (E,V,V,3) int8`,
`direction / reward / terminated / truncated: (E,)`). Host code is Python, like the reference;
the state lives in HBM and is advanced by the sm_100a kernels behind the C ABI
(multigrid_b200.engine.StepEngine). No CPU fallback: constructing an env without the CUDA
library or a CUDA device raises.

Seeding: env e of the batch behaves like a reference env whose
`env.np_random` was seeded with `seed + e` (gymnasium `reset(seed=...)`): the per-step agent
order (base.py:399) is drawn in-kernel from that PCG64 stream, bit-exactly. Layout randomness
comes from separate host generators, as in the reference (RandomMixin, utils/random.py:14-21).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _cabi, spaces
from .core.constants import Action, Color, Direction, Type
from .engine import EngineConfig, StepEngine
from .layouts import (A_COLOR, A_CC, A_CS, A_CT, A_DIR, A_TERM, A_X, A_Y, BlockedUnlockPickupLayout, EmptyLayout,
                      Layout, LockedHallwayLayout, PlaygroundLayout, RedBlueDoorsLayout)

_M64 = (1 << 64) - 1


# numpy.random.SeedSequence (bit_generator.pyx) constants
_SS_INIT_A, _SS_MULT_A = np.uint32(0x43b0d7e5), np.uint32(0x931e8875)
_SS_INIT_B, _SS_MULT_B = np.uint32(0x8b51f9dd), np.uint32(0x58f38ded)
_SS_MIX_L, _SS_MIX_R, _SS_SHIFT = np.uint32(0xca01f9dd), np.uint32(0x4973f715), np.uint32(16)


def seed_sequence_pcg64_words(entropy: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Vectorised `PCG64(SeedSequence(entropy_row))` for every row of `entropy` (uint32 [E, L], the
    32-bit words SeedSequence coerces its entropy into): (state, inc) uint64 [E,2] {lo,hi}.
    Restates numpy's SeedSequence.mix_entropy / generate_state and pcg64_srandom on uint32/uint64
    arrays (wrap-around arithmetic); checked against numpy itself in tests/test_env_api.py. Seeding
    65 536 envs takes milliseconds instead of the ~0.5 s of one SeedSequence object per env."""
    ent = np.ascontiguousarray(entropy, dtype=np.uint32)
    E, L = ent.shape
    hc = np.full(E, _SS_INIT_A, np.uint32)  # the running hash constant (same for every row)

    def hashmix(v):
        nonlocal hc
        v = v ^ hc
        hc = hc * _SS_MULT_A
        v = v * hc
        return v ^ (v >> _SS_SHIFT)

    def mix(x, y):
        r = _SS_MIX_L * x - _SS_MIX_R * y
        return r ^ (r >> _SS_SHIFT)

    with np.errstate(over="ignore"):
        zero = np.zeros(E, np.uint32)
        pool = [hashmix(ent[:, i] if i < L else zero) for i in range(4)]
        for i_src in range(4):
            for i_dst in range(4):
                if i_src != i_dst:
                    pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]))
        for i_src in range(4, L):
            for i_dst in range(4):
                pool[i_dst] = mix(pool[i_dst], hashmix(ent[:, i_src]))
        # generate_state(4, uint64) = 8 uint32 words cycling over the pool
        hb = np.full(E, _SS_INIT_B, np.uint32)
        words = []
        for i in range(8):
            v = pool[i % 4] ^ hb
            hb = hb * _SS_MULT_B
            v = v * hb
            words.append(v ^ (v >> _SS_SHIFT))
        w64 = [words[2 * i].astype(np.uint64) | (words[2 * i + 1].astype(np.uint64) << np.uint64(32)) for i in range(4)]
        # pcg64_set_seed: initstate = (w0 << 64) | w1, initseq = (w2 << 64) | w3; pcg64_srandom_r
        inc_lo = (w64[3] << np.uint64(1)) | np.uint64(1)
        inc_hi = (w64[2] << np.uint64(1)) | (w64[3] >> np.uint64(63))
        m_lo, m_hi = np.uint64(0x4385DF649FCCF645), np.uint64(0x2360ED051FC65DA4)

        def mul128(lo, hi):  # (hi:lo) * M mod 2^128 with 32-bit limbs for the 64x64 -> 128 product
            a0, a1 = lo & np.uint64(0xffffffff), lo >> np.uint64(32)
            b0, b1 = m_lo & np.uint64(0xffffffff), m_lo >> np.uint64(32)
            p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
            mid = (p00 >> np.uint64(32)) + (p01 & np.uint64(0xffffffff)) + (p10 & np.uint64(0xffffffff))
            rlo = (p00 & np.uint64(0xffffffff)) | (mid << np.uint64(32))
            rhi = p11 + (p01 >> np.uint64(32)) + (p10 >> np.uint64(32)) + (mid >> np.uint64(32))
            return rlo, rhi + lo * m_hi + hi * m_lo

        def step(lo, hi):  # state = state * M + inc
            lo, hi = mul128(lo, hi)
            nlo = lo + inc_lo
            return nlo, hi + inc_hi + (nlo < lo).astype(np.uint64)

        lo, hi = step(np.zeros(E, np.uint64), np.zeros(E, np.uint64))
        nlo = lo + w64[1]
        hi = hi + w64[0] + (nlo < lo).astype(np.uint64)
        lo, hi = step(nlo, hi)
    return np.stack([lo, hi], 1), np.stack([inc_lo, inc_hi], 1)


def _entropy_words(values: np.ndarray) -> np.ndarray:
    """Non-negative ints < 2^64 -> the uint32 words SeedSequence coerces each to (1 word below 2^32,
    else 2, little-endian); rows must agree on the count."""
    v = np.asarray(values, dtype=np.uint64)
    if (v >> np.uint64(32)).any():
        if not (v >> np.uint64(32)).all():
            raise ValueError("mixed one- and two-word seeds")
        return np.stack([(v & np.uint64(0xffffffff)).astype(np.uint32), (v >> np.uint64(32)).astype(np.uint32)], 1)
    return v.astype(np.uint32)[:, None]


def pcg64_words(seeds: Sequence[int]) -> tuple[np.ndarray, np.ndarray]:
    """(state, inc) as uint64 [E,2] {lo,hi} of `Generator(PCG64(SeedSequence(seed)))` per seed --
    what gymnasium's `reset(seed=...)` installs as `env.np_random`."""
    seeds = np.asarray(seeds)
    if len(seeds) and seeds.min() < 0:
        raise ValueError("seeds must be non-negative")  # SeedSequence rejects negative entropy
    v = seeds.astype(np.uint64)
    wide = (v >> np.uint64(32)) != 0
    st, inc = np.empty((len(v), 2), np.uint64), np.empty((len(v), 2), np.uint64)
    for sel in (wide, ~wide):
        if sel.any():
            st[sel], inc[sel] = seed_sequence_pcg64_words(_entropy_words(v[sel]))
    return st, inc


def generator_words(gen: np.random.Generator) -> tuple[tuple[int, int], tuple[int, int]]:
    d = gen.bit_generator.state["state"]
    return (d["state"] & _M64, d["state"] >> 64), (d["inc"] & _M64, d["inc"] >> 64)


def layout_generator_words(gens) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(state [K,2], inc [K,2], buf [K]) uint64 words of numpy PCG64 generators, including the buffered
    upper half of the 32-bit stream (bit 32 of buf = has_uint32, low word = uinteger)."""
    K = len(gens)
    st, inc, buf = np.empty((K, 2), np.uint64), np.empty((K, 2), np.uint64), np.empty(K, np.uint64)
    for k, g in enumerate(gens):
        d = g.bit_generator.state
        st[k] = (d["state"]["state"] & _M64, d["state"]["state"] >> 64)
        inc[k] = (d["state"]["inc"] & _M64, d["state"]["inc"] >> 64)
        buf[k] = (int(d["has_uint32"]) << 32) | int(d["uinteger"])
    return st, inc, buf


def entropy_words(count: int) -> tuple[np.ndarray, np.ndarray]:
    """Unseeded reset: any (state, odd inc) pair is a valid PCG64 stream."""
    rng = np.random.default_rng()
    st = rng.integers(0, 1 << 64, size=(count, 2), dtype=np.uint64)
    inc = rng.integers(0, 1 << 64, size=(count, 2), dtype=np.uint64)
    inc[:, 0] |= np.uint64(1)
    return st, inc


class Missions:
    """Per-env mission strings without materialising `num_envs` Python strings per step."""

    def __init__(self, table: list[str], index: torch.Tensor | None):
        self.table, self.index = table, index

    def __getitem__(self, e: int) -> str:
        return self.table[0] if self.index is None else self.table[int(self.index[e]) % len(self.table)]

    def __len__(self):
        return 0 if self.index is None else int(self.index.shape[0])

    def __repr__(self):
        return f"Missions({self.table[0]!r})" if len(self.table) == 1 else f"Missions(<{len(self.table)} strings>)"


class BatchedGrid:
    """`env.grid`: `state` is the (num_envs, width, height, 3) int8 tensor of Grid.state
    (core/grid.py:54, x-major)."""

    def __init__(self, engine: StepEngine):
        self._engine = engine
        self.width, self.height = engine.cfg.width, engine.cfg.height

    @property
    def state(self) -> torch.Tensor:
        return self._engine.grid

    def get(self, e: int, x: int, y: int):
        """Grid.get (core/grid.py:102-117) of env `e`: the WorldObj at (x, y), None for an empty cell or a position
        outside the grid. Reads one cell from the device (synchronises)."""
        if not (0 <= x < self.width and 0 <= y < self.height):
            return None
        from .core.objects import WorldObj
        return WorldObj.from_array(self._engine.grid[e, x, y].cpu().numpy())

    def set(self, e, x: int, y: int, obj) -> None:
        """Grid.set (core/grid.py:119-131): put `obj` (a WorldObj, an (type, colour, state) triple, or None = empty)
        at (x, y) of env `e` (an int, a slice, or an index tensor). Goes through the engine, which keeps the cell
        words' opaque bit, the dedup / static-grid bookkeeping and the wire palette consistent."""
        from .core.objects import WorldObj
        enc = (1, 0, 0) if obj is None else tuple(int(v) for v in (obj.encode() if isinstance(obj, WorldObj) else obj))
        grid = self._engine.grid.clone()
        grid[e, x, y] = torch.tensor(enc, dtype=grid.dtype, device=grid.device)
        self._engine.load_state(grid=grid)


class BatchedAgent:
    """`env.agents[i]`: what adapters read from `Agent` (core/agent.py:73-109), batched."""

    def __init__(self, env: "BatchedMultiGridEnv", index: int):
        self._env, self.index = env, index
        self.view_size = env.agent_view_size
        self.see_through_walls = env.see_through_walls
        self.color = Color.cycle(index + 1)[index]
        V = self.view_size
        self.observation_space = spaces.Dict({
            "image": spaces.Box(low=0, high=255, shape=(V, V, 3), dtype=np.int64),
            "direction": spaces.Discrete(len(Direction)),
            "mission": spaces.Text(max_length=256),
        })
        self.action_space = spaces.Discrete(len(Action))

    @property
    def state(self) -> torch.Tensor:
        """(E, 8) int8 packed record [dir,x,y,terminated,carry_type,carry_color,carry_state,color]."""
        return self._env.engine.agents[:, self.index]

    @property
    def pos(self) -> torch.Tensor:
        return self.state[:, A_X:A_Y + 1]

    @property
    def dir(self) -> torch.Tensor:
        return self.state[:, A_DIR]

    @property
    def terminated(self) -> torch.Tensor:
        return self.state[:, A_TERM] != 0

    @property
    def carrying(self) -> torch.Tensor:
        """(E, 3) encoding of the carried object; (1,0,0) = nothing (core/agent.py:342)."""
        return self.state[:, A_CT:A_CS + 1]

    @property
    def mission(self):
        return self._env.missions


class BatchedMultiGridEnv:
    """`num_envs` lock-stepped MultiGrid envs of one layout family on one GPU."""

    metadata = {"render_modes": []}

    def __init__(self, layout: Layout, num_envs: int = 1, device: str | torch.device = "cuda",
                 max_steps: int | None = None, see_through_walls: bool = False,
                 agent_view_size: int = 7, allow_agent_overlap: bool = True,
                 joint_reward: bool = False, success_termination_mode: str = "any",
                 failure_termination_mode: str = "all", auto_reset: bool = False,
                 pool_size: int | None = None, layout_seed: int | None = None,
                 first_env: int = 0, render_mode: str | None = None, device_layouts: bool = True,
                 stream_state: bool = False, fresh_layouts: bool = False):
        if render_mode is not None:
            raise NotImplementedError("rendering is out of scope of the batched engine")
        self.layout = layout
        self.num_envs = int(num_envs)
        self.num_agents = layout.num_agents
        self.width, self.height = layout.width, layout.height
        self.max_steps = int(max_steps or layout.max_steps)
        self.agent_view_size, self.see_through_walls = agent_view_size, see_through_walls
        self.allow_agent_overlap, self.joint_reward = allow_agent_overlap, joint_reward
        self.success_termination_mode = success_termination_mode
        self.failure_termination_mode = failure_termination_mode
        self.auto_reset = auto_reset
        self.first_env = int(first_env)  # global id of local env 0 (multi-GPU sharding)
        self.layout_seed = layout_seed
        # EmptyEnv layouts with random agent placement are generated by a CUDA kernel (bit-exact with the
        # host generator, tests/test_layouts.py); device_layouts=False forces the host path
        self.device_layouts = bool(device_layouts) and (
            (isinstance(layout, EmptyLayout) and not layout.deterministic)
            or isinstance(layout, (BlockedUnlockPickupLayout, RedBlueDoorsLayout, LockedHallwayLayout, PlaygroundLayout)))
        # fresh_layouts=True: every episode of every env gets a new _gen_grid draw from the env's own generator, like the
        # reference's reset() (one pool slot per env, regenerated on the device when the env is done); the default
        # cycles a fixed pool of `pool_size` layouts drawn at reset() (SURVEY.md section 8d's measurement protocol)
        self.fresh_layouts = bool(fresh_layouts) and auto_reset and self.device_layouts
        if fresh_layouts and not self.fresh_layouts and not layout.deterministic:
            raise ValueError("fresh_layouts needs auto_reset=True and device-side layout generation")
        self.pool_size = 1 if layout.deterministic else (
            self.num_envs if self.fresh_layouts else min(self.num_envs, pool_size or 4096))
        cfg = EngineConfig(
            width=self.width, height=self.height, num_agents=self.num_agents,
            view_size=agent_view_size, max_steps=self.max_steps,
            see_through_walls=see_through_walls, allow_agent_overlap=allow_agent_overlap,
            joint_reward=joint_reward, success_termination_mode=success_termination_mode,
            failure_termination_mode=failure_termination_mode, hook=layout.hook,
            hook_param=getattr(layout, "hook_param", 0), auto_reset=auto_reset,
            stream_state=stream_state)  # (cache policy only, see EngineConfig.stream_state)
        self.engine = StepEngine(cfg, self.num_envs, device)
        self.device = self.engine.device
        self.grid = BatchedGrid(self.engine)
        self.agents = [BatchedAgent(self, i) for i in range(self.num_agents)]
        self.observation_space = spaces.Dict({i: a.observation_space for i, a in enumerate(self.agents)})
        self.action_space = spaces.Dict({i: a.action_space for i, a in enumerate(self.agents)})
        self.missions = Missions([layout.mission], None)
        self._actions = torch.full((self.num_envs, self.num_agents), -1, dtype=torch.int8,
                                   device=self.device)
        self._needs_reset = True
        self._step_views = None

    # -- reference attributes -------------------------------------------------------------------
    @property
    def unwrapped(self):
        return self

    @property
    def agent_states(self) -> torch.Tensor:
        """(E, n, 8) int8 packed AgentState (core/agent.py:222-232 without the constant TYPE)."""
        return self.engine.agents

    @property
    def step_count(self) -> torch.Tensor:
        return self.engine.step_count

    def is_done(self) -> torch.Tensor:
        """(E,) bool: base.py:534-539."""
        term = self.engine.agents[:, :, A_TERM] != 0
        return (self.engine.step_count >= self.max_steps) | term.all(dim=1)

    # -- reset ----------------------------------------------------------------------------------
    def _seeds(self, seed) -> np.ndarray | None:
        if seed is None:
            return None
        if np.ndim(seed) == 0:
            return int(seed) + self.first_env + np.arange(self.num_envs, dtype=np.int64)
        seeds = np.asarray(seed, dtype=np.int64)
        if seeds.shape != (self.num_envs,):
            raise ValueError(f"seed must be an int or {self.num_envs} ints")
        return seeds

    def reset(self, seed=None, options: dict | None = None):
        """-> (obs, infos). `seed`: int (env e gets seed + first_env + e) or one int per env.
        `options['layout_rngs']`: one numpy Generator per pool entry (parity tests)."""
        self.engine.check_status()
        options = options or {}
        E, K = self.num_envs, self.pool_size
        seeds = self._seeds(seed)
        st, inc = entropy_words(E) if seeds is None else pcg64_words(seeds)
        layout_rngs = options.get("layout_rngs")

        def layout_rng(k):
            if layout_rngs is not None:
                return layout_rngs[k]
            if self.layout_seed is not None:
                return np.random.default_rng([int(self.layout_seed), self.first_env + k])
            return np.random.default_rng()

        if self.device_layouts:
            gens = None
            if layout_rngs is not None:
                gens = [layout_rngs[k] for k in range(K)]
                lst, linc, lbuf = layout_generator_words(gens)
            elif self.layout_seed is not None and 0 <= int(self.layout_seed) < 2 ** 32:
                # == default_rng([layout_seed, first_env + k]) for every k, without K generator objects
                ent = np.stack([np.full(K, int(self.layout_seed), np.uint32),
                                (self.first_env + np.arange(K)).astype(np.uint32)], 1)
                lst, linc = seed_sequence_pcg64_words(ent)
                lbuf = np.zeros(K, np.uint64)
            elif self.layout_seed is not None:
                lst, linc, lbuf = layout_generator_words([layout_rng(k) for k in range(K)])
            else:
                lst, linc = entropy_words(K)
                lbuf = np.zeros(K, np.uint64)
            table = [self.layout.mission]
            if isinstance(self.layout, BlockedUnlockPickupLayout):
                # env k's own order stream gives layout k's door height and is advanced by that draw
                ost, box_color, lst, lbuf = self.engine.gen_layout_pool_bup(
                    self.layout.room_size, lst, linc, lbuf, st[:K], inc[:K])
                st[:K] = ost
                names = [c.value for c in Color]
                table = [f"pick up the {names[int(c)]} box" for c in box_color]  # blockedunlockpickup.py:139-140
            elif isinstance(self.layout, PlaygroundLayout):
                lo = self.layout
                ost, _, lst, lbuf = self.engine.gen_layout_pool_playground(
                    lo.room_size, lo.num_rows, lo.num_cols, lst, linc, lbuf, st[:K], inc[:K])
                st[:K] = ost
            elif isinstance(self.layout, LockedHallwayLayout):
                lo = self.layout
                lst, lbuf = self.engine.gen_layout_pool_locked_hallway(
                    lo.num_rooms, lo.room_size, lo.max_hallway_keys, lo.max_keys_per_room, lst, linc, lbuf)
            elif isinstance(self.layout, RedBlueDoorsLayout):
                lst, lbuf = self.engine.gen_layout_pool_red_blue_doors(self.layout.size, lst, linc, lbuf)
            else:
                lst, lbuf = self.engine.gen_layout_pool_empty_random(lst, linc, lbuf)
            if gens is not None:  # the caller's generators advance as the reference's would
                for k, g in enumerate(gens):
                    d = g.bit_generator.state
                    d["state"]["state"] = int(lst[k, 0]) | (int(lst[k, 1]) << 64)
                    d["has_uint32"], d["uinteger"] = int(lbuf[k] >> np.uint64(32)) & 1, int(lbuf[k] & np.uint64(0xffffffff))
                    g.bit_generator.state = d
            idx = np.arange(E, dtype=np.int32) % K
            self.engine.load_state(layout_idx=idx, pcg_state=st, pcg_inc=inc)
            self.engine.reset_from_pool()
            self.missions = Missions(table, None if len(set(table)) == 1 else self.engine.layout_idx)
            if self.fresh_layouts:
                self.engine.enable_fresh_layouts()
                if isinstance(self.layout, BlockedUnlockPickupLayout):  # the mission names the slot's box colour
                    names = [c.value for c in Color]
                    self.missions = Missions([f"pick up the {nm} box" for nm in names], self.engine._pool_gen[3])
            self._needs_reset = False
            return self._obs(self.engine.gen_obs()), defaultdict(dict)

        grids, agents, infos = [], [], []
        for k in range(K):
            lrng = layout_rng(k)
            if self.layout.deterministic:
                orng = None
            else:  # env k's own order stream: reset-time draws (door positions) advance it
                orng = np.random.Generator(np.random.PCG64())
                s = orng.bit_generator.state
                s["state"] = {"state": int(st[k, 0]) | (int(st[k, 1]) << 64),
                              "inc": int(inc[k, 0]) | (int(inc[k, 1]) << 64)}
                orng.bit_generator.state = s
            g, a, info = self.layout.generate(lrng, orng)
            if orng is not None:
                st[k], _ = generator_words(orng)
            grids.append(g)
            agents.append(a)
            infos.append(info)
        pool_grid, pool_agents = np.stack(grids), np.stack(agents)
        self.engine.set_layout_pool(pool_grid, pool_agents)
        table = [info.get("mission", self.layout.mission) for info in infos]
        idx = np.arange(E, dtype=np.int32) % K
        self.engine.load_state(layout_idx=idx, pcg_state=st, pcg_inc=inc)
        self.engine.reset_from_pool()
        self.missions = Missions(table, None if len(set(table)) == 1 else self.engine.layout_idx)
        self._needs_reset = False
        return self._obs(self.engine.gen_obs()), defaultdict(dict)

    # -- step -----------------------------------------------------------------------------------
    def _obs(self, image: torch.Tensor) -> dict:
        direction = self.engine.direction
        return {i: {"image": image[:, i], "direction": direction[:, i], "mission": self.missions}
                for i in range(self.num_agents)}

    def _action_tensor(self, actions) -> torch.Tensor:
        E, n = self.num_envs, self.num_agents
        if isinstance(actions, torch.Tensor):
            if actions.shape != (E, n):
                raise ValueError(f"actions tensor must have shape ({E}, {n})")
            if actions.dtype == torch.int8 and actions.device == self.device and actions.is_contiguous():
                return actions
            self.engine._chain_armed = None  # staged below: the launch must not rely on a chained predecessor
            if actions.dtype != torch.int8 and not actions.is_floating_point():
                # values that do not fit int8 must not wrap into valid actions: anything outside -1..6 becomes the
                # invalid code 7, which the kernel flags (the reference raises ValueError, base.py:473-474)
                actions = torch.where((actions < -1) | (actions > 6), 7, actions)
            self._actions.copy_(actions)
            return self._actions
        self.engine._chain_armed = None
        if isinstance(actions, np.ndarray):
            a = np.asarray(actions)
            if a.dtype != np.int8:
                a = np.where((a < -1) | (a > 6), 7, a)
            self._actions.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.int8)).reshape(E, n))
            return self._actions
        # dict {agent_id: action}; ids missing from the dict do not act (base.py:403-404)
        if len(actions) == n and all(isinstance(actions.get(i), torch.Tensor) and actions[i].device == self.device
                                     and actions[i].shape == (E,) for i in range(n)):
            cols = [actions[i] for i in range(n)]  # one kernel instead of a fill and n column copies
            if all(c.dtype == torch.int8 for c in cols):
                torch.stack(cols, dim=1, out=self._actions)
            else:
                self._actions.copy_(torch.stack(cols, dim=1))
            return self._actions
        self._actions.fill_(-1)
        for i, a in actions.items():
            if not 0 <= int(i) < n:
                continue
            if isinstance(a, torch.Tensor):
                self._actions[:, i] = a.to(self.device, torch.int8)
            elif np.ndim(a) == 0:
                if not 0 <= int(a) < len(Action):
                    raise ValueError(f"Unknown action: {a}")  # base.py:473-474
                self._actions[:, i] = int(a)
            else:
                self._actions[:, i] = torch.as_tensor(np.asarray(a, dtype=np.int8), device=self.device)
        return self._actions

    def step(self, actions, chained: bool = False):
        """-> (obs, rewards, terminations, truncations, infos), dicts keyed by agent index whose
        values are batched device tensors (views of engine buffers, overwritten by the next call).
        Asynchronous: nothing here synchronises with the GPU. Out-of-range actions inside tensors
        are flagged on the device and raise ValueError at the next `reset()` / `check()`."""
        if self._needs_reset:
            raise RuntimeError("call reset() before step()")
        # chained=True: StepEngine.step's MG_FLAG_CHAINED (open-loop action tapes only, see its docstring)
        image, reward, terminated, truncated = self.engine.step(self._action_tensor(actions), chained=chained)
        if self._step_views is None or self._step_views[0] is not image:
            # every item is a view of a fixed engine buffer: build the per-agent views once (some 20 tensor
            # ops, twice the cost of the kernel launch) and hand out fresh dicts around them every step
            term_b, trunc_b = terminated.view(torch.bool), truncated.view(torch.bool)
            n, direction = self.num_agents, self.engine.direction
            self._step_views = (image,
                                [(image[:, i], direction[:, i]) for i in range(n)],
                                [reward[:, i] for i in range(n)],
                                [term_b[:, i] for i in range(n)], trunc_b)
        _, obs_v, rew_v, term_v, trunc_b = self._step_views
        missions = self.missions
        return ({i: {"image": v[0], "direction": v[1], "mission": missions} for i, v in enumerate(obs_v)},
                dict(enumerate(rew_v)), dict(enumerate(term_v)), dict.fromkeys(range(len(rew_v)), trunc_b),
                defaultdict(dict))

    def reset_where(self, mask):
        """Reset only the envs selected by `mask` ((E,) bool tensor / array): each takes the next layout of
        the pool (what `auto_reset=True` does inside the kernel, driven from outside, as RLlib does with its
        sub-envs). Returns the observations of the whole batch after the reset."""
        if self._needs_reset:
            raise RuntimeError("call reset() before reset_where()")
        if not isinstance(mask, torch.Tensor):
            mask = torch.as_tensor(np.asarray(mask, dtype=bool))
        self.engine.reset_where(mask.to(self.device))
        return self._obs(self.engine.gen_obs())

    def check(self) -> None:
        """Synchronise and raise ValueError if a kernel saw an unknown action (base.py:473-474)."""
        self.engine.check_status()

    def close(self) -> None:
        pass
