#!/usr/bin/env python
"""Generate the committed golden fixtures by EXECUTING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE. Run in the build container only (it needs /root/reference, which does
not exist on the GPU box):

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py

The reference (ini/multigrid @ /root/reference) imports three third-party packages that this
image lacks (gymnasium, aenum, pygame); `tests/golden/shims/` provides minimal stand-ins so the
reference's own files run unmodified.

What is recorded, per case (one .npz each under tests/golden/):
  * the post-reset state of B reference envs: `grid.state` (W,H,3), `agent_states` (n,9) and the
    128-bit PCG64 (state, inc) of `env.np_random` AFTER reset (base.py:399 draws the per-step
    agent order from it; RoomGrid.add_door already consumed draws during reset, roomgrid.py:324),
  * a random action tape (T,B,n) with -1 = "agent id absent from the action dict",
  * after every `env.step`: images, directions, rewards, terminations, truncations, plus the
    full grid / agent state (so a mismatch can be localised to the transition or the observation).

Deterministic-oracle recipe: the layout RNG captured by RandomMixin at
construction (base.py:143) and the gymnasium RNG are both replaced with seeded generators
before `reset()` is called with no seed.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("MULTIGRID_REFERENCE", "/root/reference")
sys.path.insert(0, REFERENCE)
sys.path.insert(0, os.path.join(HERE, "shims"))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

import numpy as np  # noqa: E402
import gymnasium as gym  # noqa: E402  (shim)
import multigrid.envs  # noqa: E402,F401  (reference; registers the ids)
from multigrid.base import MultiGridEnv  # noqa: E402
from multigrid.core.constants import Color, Direction, Type  # noqa: E402
from multigrid.core.grid import Grid  # noqa: E402
from multigrid.core.world_object import Ball, Box, Door, Floor, Goal, Key, Lava, Wall  # noqa: E402
from multigrid.utils.obs import gen_obs_grid_encoding  # noqa: E402

M64 = (1 << 64) - 1


class SoupEnv(MultiGridEnv):
    """Dense random object soup: makes every rule of handle_actions fire under random actions.

    Not a reference env: a subclass that only implements `_gen_grid`; the transition and the
    observation code exercised is 100% the reference's (base.py / utils/obs.py).
    """

    def __init__(self, width=9, height=7, density=0.45, **kwargs):
        self.density = density
        super().__init__(mission_space="soup", width=width, height=height, **kwargs)

    def _gen_grid(self, width, height):
        self.grid = Grid(width, height)
        self.grid.wall_rect(0, 0, width, height)
        colors = list(Color)
        for x in range(1, width - 1):
            for y in range(1, height - 1):
                if self._rand_float(0, 1) > self.density:
                    continue
                kind = self._rand_int(0, 10)
                color = colors[self._rand_int(0, len(colors))]
                if kind == 0:
                    obj = Wall()
                elif kind in (1, 2):
                    state = self._rand_int(0, 3)
                    obj = Door(color, is_open=(state == 0), is_locked=(state == 2))
                elif kind in (3, 4):
                    obj = Key(color)
                elif kind == 5:
                    obj = Ball(color)
                elif kind == 6:
                    obj = Box(color)
                elif kind == 7:
                    obj = Goal()
                elif kind == 8:
                    obj = Lava()
                else:
                    obj = Floor(color)
                self.grid.set(x, y, obj)
        for agent in self.agents:
            self.place_agent(agent)
            # some agents start with a key in hand so locked doors get opened
            if self._rand_int(0, 3) == 0:
                agent.state.carrying = Key(colors[self._rand_int(0, len(colors))])


gym.register(id="Golden-Soup-v0", entry_point=SoupEnv, kwargs={})


def pcg_words(env):
    st = env.np_random.bit_generator.state["state"]
    s, inc = int(st["state"]), int(st["inc"])
    return [s & M64, s >> 64], [inc & M64, inc >> 64]


def make_env(env_id, kwargs, layout_seed, order_seed):
    env = gym.make(env_id, **kwargs)
    env._RandomMixin__np_random = np.random.default_rng(layout_seed)
    env._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(order_seed)))
    return env


def collect_obs(obs, n):
    img = np.stack([obs[i]["image"] for i in range(n)])
    direction = np.array([int(obs[i]["direction"]) for i in range(n)])
    return img, direction


def bup_teleport(env):
    """Put agent 0 next to the box, facing it, so the BlockedUnlockPickup success hook
    (envs/blockedunlockpickup.py:166-175) fires under random actions. State injection only."""
    bx, by = env.obj.cur_pos
    for d, (dx, dy) in enumerate([(1, 0), (0, 1), (-1, 0), (0, -1)]):
        x, y = bx - dx, by - dy
        if env.grid.get(x, y) is None:
            env.agents[0].state.pos = (x, y)
            env.agents[0].state.dir = d
            return
    raise RuntimeError("no free cell next to the box")


def run_case(name, env_id, kwargs, B, T, seed, p_absent=0.0, action_p=None, auto_reset=False,
             tweak=None, save=True):
    """Roll B reference envs for T steps, recording everything. See module docstring.
    save=False returns the record instead of writing tests/golden/<name>.npz (tests/test_live_reference.py)."""
    rng = np.random.default_rng(seed)
    envs, rec = [], {}
    init_grid, init_agents, pcg_state, pcg_inc, obs0, dir0 = [], [], [], [], [], []
    for b in range(B):
        env = make_env(env_id, kwargs, layout_seed=seed * 1000 + b, order_seed=seed * 7919 + b)
        obs, _ = env.reset()
        if tweak is not None:
            tweak(env)
            obs = env.gen_obs()
        envs.append(env)
        n = env.num_agents
        init_grid.append(env.grid.state.copy())
        init_agents.append(np.asarray(env.agent_states).copy())
        s, inc = pcg_words(env)
        pcg_state.append(s)
        pcg_inc.append(inc)
        img, d = collect_obs(obs, n)
        obs0.append(img)
        dir0.append(d)
    env0 = envs[0]
    n, V = env0.num_agents, env0.agents[0].view_size
    W, H = env0.width, env0.height

    actions = rng.choice(7, size=(T, B, n), p=action_p).astype(np.int8)
    if p_absent > 0:
        actions[rng.random((T, B, n)) < p_absent] = -1

    obs_t = np.zeros((T, B, n, V, V, 3), np.int8)
    dir_t = np.zeros((T, B, n), np.int8)
    rew_t = np.zeros((T, B, n), np.float64)
    term_t = np.zeros((T, B, n), np.uint8)
    trunc_t = np.zeros((T, B), np.uint8)
    grid_t = np.zeros((T, B, W, H, 3), np.int8)
    agents_t = np.zeros((T, B, n, 9), np.int8)
    step_count_t = np.zeros((T, B), np.int32)
    # layouts used by auto-reset: episode j of env b -> pool slot b*J + j
    pool_grid = [[g] for g in init_grid]
    pool_agents = [[a] for a in init_agents]
    done_prev = [False] * B
    n_events = dict(resets=0, rewards=0, terms=0)

    for t in range(T):
        for b, env in enumerate(envs):
            if auto_reset and done_prev[b]:
                # "next-step" auto-reset: this call resets instead of stepping; the action-order
                # stream is NOT advanced by a reset in the batched engine, so restore it.
                saved = env.np_random.bit_generator.state
                obs, _ = env.reset()
                env.np_random.bit_generator.state = saved
                pool_grid[b].append(env.grid.state.copy())
                pool_agents[b].append(np.asarray(env.agent_states).copy())
                rew = {i: 0.0 for i in range(n)}
                term = {i: False for i in range(n)}
                trunc = {i: False for i in range(n)}
                n_events["resets"] += 1
            else:
                act = {i: int(actions[t, b, i]) for i in range(n) if actions[t, b, i] >= 0}
                obs, rew, term, trunc, _ = env.step(act)
            img, d = collect_obs(obs, n)
            obs_t[t, b], dir_t[t, b] = img, d
            rew_t[t, b] = [float(rew[i]) for i in range(n)]
            term_t[t, b] = [bool(term[i]) for i in range(n)]
            trunc_t[t, b] = bool(trunc[0])
            grid_t[t, b] = env.grid.state
            agents_t[t, b] = np.asarray(env.agent_states)
            step_count_t[t, b] = env.step_count
            done_prev[b] = bool(all(term_t[t, b]) or trunc_t[t, b])
            n_events["rewards"] += int((rew_t[t, b] != 0).any())
            n_events["terms"] += int(term_t[t, b].any())

    J = max(len(p) for p in pool_grid)
    pg = np.zeros((B * J, W, H, 3), np.int8)
    pa = np.zeros((B * J, n, 9), np.int8)
    for b in range(B):
        for j in range(J):
            jj = min(j, len(pool_grid[b]) - 1)
            pg[b * J + j] = pool_grid[b][jj]
            pa[b * J + j] = pool_agents[b][jj]

    meta = dict(
        env_id=env_id, kwargs=repr(kwargs), B=B, T=T, n=n, V=V, W=W, H=H,
        max_steps=env0.max_steps,
        see_through_walls=int(env0.agents[0].see_through_walls),
        allow_agent_overlap=int(env0.allow_agent_overlap),
        joint_reward=int(env0.joint_reward),
        success_any=int(env0.success_termination_mode == "any"),
        failure_any=int(env0.failure_termination_mode == "any"),
        hook={"BlockedUnlockPickupEnv": 1, "RedBlueDoorsEnv": 2, "LockedHallwayEnv": 3}.get(
            type(env0).__name__, 0),
        hook_param=getattr(env0, "num_rooms", 0),
        auto_reset=int(auto_reset), pool_J=J,
    )
    rec = dict(
        init_grid=np.stack(init_grid).astype(np.int8),
        init_agents=np.stack(init_agents).astype(np.int8),
        pcg_state=np.array(pcg_state, dtype=np.uint64),
        pcg_inc=np.array(pcg_inc, dtype=np.uint64),
        obs0=np.stack(obs0).astype(np.int8), dir0=np.stack(dir0).astype(np.int8),
        actions=actions, obs=obs_t, direction=dir_t, reward=rew_t, terminated=term_t,
        truncated=trunc_t, grid=grid_t, agents=agents_t, step_count=step_count_t,
        pool_grid=pg, pool_agents=pa,
        **{f"meta_{k}": np.array(v) for k, v in meta.items()},
    )
    if not save:
        return rec, meta
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: B={B} T={T} n={n} V={V} {W}x{H} events={n_events} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def one_hot_kat(name, seed=3, cases=24):
    """OneHotObsWrapper.one_hot (wrappers.py:158-190, numba) on observation-shaped arrays."""
    from multigrid.wrappers import OneHotObsWrapper
    rng = np.random.default_rng(seed)
    dim_sizes = np.array([11, 6, 4])
    xs, outs = [], []
    for c in range(cases):
        V = (3, 5, 7, 9)[c % 4]
        x = np.stack([rng.integers(0, 11, (V, V)), rng.integers(0, 6, (V, V)), rng.integers(0, 4, (V, V))], -1)
        out = OneHotObsWrapper.one_hot(x.astype(np.int64), dim_sizes)
        pad = np.zeros((9, 9, 3), np.int8); pad[:V, :V] = x
        po = np.zeros((9, 9, 21), np.uint8); po[:V, :V] = out
        xs.append(pad); outs.append(po)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), x=np.stack(xs), out=np.stack(outs),
                        V=np.array([(3, 5, 7, 9)[c % 4] for c in range(cases)]))
    print(f"{name}: ok")


def full_obs_kat(name, seed=9):
    """FullyObsWrapper (wrappers.py:17-58) images along short reference rollouts, with the state they
    were made from (one agent is terminated by hand so that terminated agents are covered)."""
    from multigrid.wrappers import FullyObsWrapper
    rng = np.random.default_rng(seed)
    grids, agents, imgs, dims = [], [], [], []
    for env_id, kw in (("MultiGrid-Empty-8x8-v0", dict(agents=3)),
                       ("MultiGrid-BlockedUnlockPickup-v0", dict(agents=2)),
                       ("MultiGrid-Empty-Random-6x6-v0", dict(agents=4))):
        env = FullyObsWrapper(make_env(env_id, kw, layout_seed=seed, order_seed=seed + 1))
        obs, _ = env.reset()
        base = env.unwrapped
        for t in range(25):
            if t == 10:
                base.agents[0].state.terminated = True
            act = {i: int(rng.integers(0, 7)) for i in range(base.num_agents)}
            obs, *_ = env.step(act)
            W, H = base.width, base.height
            g = np.zeros((16, 16, 3), np.int8); g[:W, :H] = base.grid.state
            im = np.zeros((16, 16, 3), np.int8); im[:W, :H] = obs[0]["image"]
            a = np.zeros((4, 9), np.int8); a[:base.num_agents] = np.asarray(base.agent_states)
            grids.append(g); imgs.append(im); agents.append(a); dims.append((W, H, base.num_agents))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), grid=np.stack(grids), agents=np.stack(agents),
                        img=np.stack(imgs), dims=np.array(dims))
    print(f"{name}: {len(dims)} states")


def random_obs_cases(name, seed, cases=400):
    """Pure-function golden vectors for utils/obs.py:66-102 on injected random states."""
    rng = np.random.default_rng(seed)
    groups = {}
    for c in range(cases):
        W, H = int(rng.integers(3, 14)), int(rng.integers(3, 14))
        n = int(rng.integers(1, 6))
        V = int(rng.choice([3, 5, 7, 9]))
        stw = bool(rng.random() < 0.2)
        grid = np.zeros((W, H, 3), dtype=np.int_)
        grid[..., 0] = 1
        r = rng.random((W, H))
        for x in range(W):
            for y in range(H):
                if r[x, y] < 0.25:
                    grid[x, y] = (2, 5, 0)
                elif r[x, y] < 0.35:
                    grid[x, y] = (4, rng.integers(6), rng.integers(3))
                elif r[x, y] < 0.45:
                    t = int(rng.integers(5, 10))
                    grid[x, y] = (t, rng.integers(6), 0)
                elif r[x, y] < 0.50:
                    grid[x, y] = (3, rng.integers(6), 0)
        agents = np.zeros((n, 9), dtype=np.int_)
        agents[:, 0] = 10
        agents[:, 1] = np.arange(n) % 6
        agents[:, 2] = rng.integers(0, 4, n)
        agents[:, 3] = rng.integers(0, W, n)
        agents[:, 4] = rng.integers(0, H, n)
        agents[:, 5] = rng.random(n) < 0.2
        agents[:, 6] = 1
        for k in range(n):
            if rng.random() < 0.3:
                agents[k, 6:9] = (rng.integers(5, 8), rng.integers(6), 0)
        obs = gen_obs_grid_encoding(grid, agents, V, stw)
        groups.setdefault("W", []).append(W)
        groups.setdefault("H", []).append(H)
        groups.setdefault("n", []).append(n)
        groups.setdefault("V", []).append(V)
        groups.setdefault("stw", []).append(int(stw))
        # ragged -> pad into fixed maxima so one npz holds everything
        gp = np.zeros((13, 13, 3), np.int8); gp[:W, :H] = grid
        ap = np.zeros((5, 9), np.int8); ap[:n] = agents
        op = np.zeros((5, 9, 9, 3), np.int8); op[:n, :V, :V] = obs
        groups.setdefault("grid", []).append(gp)
        groups.setdefault("agents", []).append(ap)
        groups.setdefault("obs", []).append(op)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **{k: np.array(v) for k, v in groups.items()})
    print(f"{name}: {cases} cases -> {os.path.getsize(path) / 1024:.0f} KiB")


def pcg_kat(name):
    """Known-answer vectors for numpy Generator(PCG64).random() (call site base.py:399)."""
    seeds, states, incs, draws = [], [], [], []
    for seed in (0, 1, 7, 123, 2**31 - 1):
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        st = g.bit_generator.state["state"]
        seeds.append(seed)
        states.append([st["state"] & M64, st["state"] >> 64])
        incs.append([st["inc"] & M64, st["inc"] >> 64])
        draws.append(g.random(size=16))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), seeds=np.array(seeds),
                        state=np.array(states, dtype=np.uint64),
                        inc=np.array(incs, dtype=np.uint64), draws=np.array(draws))
    print(f"{name}: ok")


FWD_HEAVY = [0.15, 0.15, 0.30, 0.12, 0.10, 0.15, 0.03]
PICKUP_HEAVY = [0.10, 0.10, 0.15, 0.40, 0.10, 0.10, 0.05]

def rbd_open_doors(env):
    """Open the red door (and for odd envs leave it closed) and put agent 0 in front of the blue
    door, so that success, failure and the re-closing of the blue door (envs/redbluedoors.py:170-187)
    all occur under random actions. State injection only."""
    rx, ry = [(x, y) for x in range(env.width) for y in range(env.height) if env.grid.get(x, y) is env.red_door][0]
    if env.np_random.random() < 0.5:
        env.red_door.is_open = True
        env.grid.update(rx, ry)  # keep grid.state in sync with the object, as Door.toggle does
    bx, by = [(x, y) for x in range(env.width) for y in range(env.height) if env.grid.get(x, y) is env.blue_door][0]
    env.agents[0].state.pos = (bx - 1, by)
    env.agents[0].state.dir = 0
    if env.num_agents > 1:
        env.agents[1].state.pos = (rx + 1, ry)
        env.agents[1].state.dir = 2


def lh_keys_at_doors(env):
    """Hand agent k the key of the k-th door and put it in the hallway cell in front of that door,
    facing it, so that unlock rewards (joint and not), repeated toggles of an already unlocked
    door and the all-doors-unlocked termination (envs/locked_hallway.py:203-227) occur under
    random actions. State injection only."""
    from multigrid.core.world_object import Door as _Door, Key as _Key
    doors = [(x, y, env.grid.get(x, y)) for x in range(env.width) for y in range(env.height)
             if isinstance(env.grid.get(x, y), _Door)]
    hall_x0 = env.get_room(1, 0).top[0]
    for k, agent in enumerate(env.agents):
        x, y, door = doors[k % len(doors)]
        left_door = x == hall_x0  # door in the hallway's left wall: stand right of it, face left
        agent.state.pos = (x + 1, y) if left_door else (x - 1, y)
        agent.state.dir = 2 if left_door else 0
        agent.state.carrying = _Key(door.color)


TOGGLE_HEAVY = [0.10, 0.10, 0.15, 0.05, 0.05, 0.50, 0.05]

if __name__ == "__main__":
    only = set(sys.argv[1:])
    _run_case, _pcg_kat, _random_obs_cases = run_case, pcg_kat, random_obs_cases

    def run_case(name, *a, **k):  # noqa: F811  (`python make_golden.py NAME...` regenerates only those)
        if not only or name in only:
            _run_case(name, *a, **k)

    def pcg_kat(name, *a, **k):  # noqa: F811
        if not only or name in only:
            _pcg_kat(name, *a, **k)

    def random_obs_cases(name, *a, **k):  # noqa: F811
        if not only or name in only:
            _random_obs_cases(name, *a, **k)

    if not only or "one_hot_kat" in only:
        one_hot_kat("one_hot_kat")
    if not only or "full_obs_kat" in only:
        full_obs_kat("full_obs_kat")

    pcg_kat("pcg64_kat")
    random_obs_cases("obs_random", seed=11)
    # BASELINE.json configs[0..3]
    run_case("empty8_n2", "MultiGrid-Empty-8x8-v0", dict(agents=2), B=4, T=300, seed=1)
    run_case("empty8_n4", "MultiGrid-Empty-8x8-v0", dict(agents=4), B=4, T=300, seed=2)
    run_case("bup_n2", "MultiGrid-BlockedUnlockPickup-v0", dict(agents=2), B=6, T=400, seed=3,
             action_p=FWD_HEAVY)
    run_case("bup_n2_teleport", "MultiGrid-BlockedUnlockPickup-v0", dict(agents=2), B=6, T=60,
             seed=33, action_p=PICKUP_HEAVY, tweak=bup_teleport)
    run_case("bup_n3_teleport_nojoint", "MultiGrid-BlockedUnlockPickup-v0",
             dict(agents=3, joint_reward=False, allow_agent_overlap=False), B=4, T=60,
             seed=34, action_p=PICKUP_HEAVY, tweak=bup_teleport)
    run_case("empty16_n8_v9", "MultiGrid-Empty-16x16-v0", dict(agents=8, agent_view_size=9),
             B=2, T=200, seed=4)
    # flag coverage
    run_case("empty6r_n3_nooverlap_all", "MultiGrid-Empty-Random-6x6-v0",
             dict(agents=3, allow_agent_overlap=False, success_termination_mode="all",
                  agent_start_dir=None), B=6, T=200, seed=5, p_absent=0.1)
    run_case("empty5_n1", "MultiGrid-Empty-5x5-v0", dict(agents=1, agent_view_size=3),
             B=4, T=120, seed=6)
    run_case("empty8_n4_joint_stw", "MultiGrid-Empty-8x8-v0",
             dict(agents=4, joint_reward=True, see_through_walls=True, agent_view_size=5),
             B=4, T=300, seed=7)
    run_case("playground_n3", "MultiGrid-Playground-v0", dict(agents=3), B=3, T=150, seed=8,
             action_p=FWD_HEAVY)
    for i, kw in enumerate([
        dict(agents=3, failure_termination_mode="all", success_termination_mode="all"),
        dict(agents=4, failure_termination_mode="any", success_termination_mode="any",
             joint_reward=True, allow_agent_overlap=False),
        dict(agents=2, failure_termination_mode="all", success_termination_mode="any",
             joint_reward=True, width=7, height=12, agent_view_size=9, max_steps=60),
    ]):
        run_case(f"soup_{i}", "Golden-Soup-v0", kw, B=12, T=120, seed=20 + i,
                 action_p=FWD_HEAVY, p_absent=0.05)
    # auto-reset ("next-step" mode) on reference envs
    run_case("empty8_n4_autoreset", "MultiGrid-Empty-8x8-v0", dict(agents=4, max_steps=40),
             B=4, T=200, seed=30, auto_reset=True)
    run_case("soup_autoreset", "Golden-Soup-v0",
             dict(agents=3, max_steps=25, failure_termination_mode="any"),
             B=6, T=150, seed=31, action_p=FWD_HEAVY, auto_reset=True)
    run_case("bup_n2_autoreset", "MultiGrid-BlockedUnlockPickup-v0", dict(agents=2, max_steps=30),
             B=4, T=120, seed=32, action_p=FWD_HEAVY, auto_reset=True)
    # several episodes per env of the random-layout ids: every reset draws a NEW layout from the env's generator
    run_case("empty6r_n3_autoreset", "MultiGrid-Empty-Random-6x6-v0", dict(agents=3, max_steps=18),
             B=5, T=100, seed=35, auto_reset=True)
    run_case("playground_n2_autoreset", "MultiGrid-Playground-v0", dict(agents=2, max_steps=20),
             B=3, T=90, seed=36, action_p=FWD_HEAVY, auto_reset=True)
    # RedBlueDoors post-hook (success / failure / blue door closed again), SURVEY.md section 8f N3
    run_case("rbd_n2", "MultiGrid-RedBlueDoors-8x8-v0", dict(agents=2), B=8, T=150, seed=40,
             action_p=TOGGLE_HEAVY, tweak=rbd_open_doors)
    run_case("rbd6_n3_all", "MultiGrid-RedBlueDoors-6x6-v0",
             dict(agents=3, success_termination_mode="all", failure_termination_mode="all",
                  joint_reward=False), B=8, T=150, seed=41, action_p=TOGGLE_HEAVY, tweak=rbd_open_doors)
    run_case("rbd_n2_autoreset", "MultiGrid-RedBlueDoors-6x6-v0", dict(agents=2, max_steps=40),
             B=4, T=160, seed=42, action_p=TOGGLE_HEAVY, auto_reset=True)
    # LockedHallway post-hook (first-unlock rewards that ADD UP, termination only in the returned dict)
    run_case("lh2_n2", "MultiGrid-LockedHallway-2Rooms-v0", dict(agents=2), B=6, T=80, seed=50,
             action_p=TOGGLE_HEAVY, tweak=lh_keys_at_doors)
    run_case("lh4_n3_nojoint", "MultiGrid-LockedHallway-4Rooms-v0", dict(agents=3, joint_reward=False),
             B=6, T=80, seed=51, action_p=TOGGLE_HEAVY, tweak=lh_keys_at_doors)
    run_case("lh6_n4", "MultiGrid-LockedHallway-6Rooms-v0", dict(agents=4), B=4, T=100, seed=52,
             action_p=FWD_HEAVY)
    run_case("lh2_n2_autoreset", "MultiGrid-LockedHallway-2Rooms-v0", dict(agents=2, max_steps=30),
             B=4, T=120, seed=53, action_p=TOGGLE_HEAVY, auto_reset=True, tweak=lh_keys_at_doors)
