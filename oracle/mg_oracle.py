"""CPU ORACLE (numpy / pure Python) for the MultiGrid step/observe hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE. Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu-baseline leg may import this module; `multigrid_b200/` never does.

It restates, on the packed int8 state the engine keeps in HBM, the algorithm of the reference
(ini/multigrid, `/root/reference`). Every function cites the reference lines it follows.

Parity status: PINNED. `tests/test_oracle_golden.py` checks this restatement against fixtures
recorded by executing the unmodified reference (`tests/golden/make_golden.py`): 16 rollout cases
(all registered env families used by BASELINE.json + flag/termination-mode coverage + a dense
"soup" of every object type) and 400 random injected observation states, plus numpy PCG64
known-answer vectors. The reference itself ships no tests or golden vectors.

Packed state (per env):
    grid        int8 (W, H, 3)   x-major like `Grid.state` (core/grid.py:54): (type, color, state)
    agents      int8 (n, 8)      [dir, x, y, terminated, carry_type, carry_color, carry_state, color]
                                 = AgentState (core/agent.py:222-232) minus the constant TYPE=10
    step_count  int32
    pcg_state   uint64 (2,)      [lo, hi] of numpy PCG64's 128-bit state (env.np_random)
    pcg_inc     uint64 (2,)      [lo, hi] of its 128-bit increment
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# --- encodings: core/constants.py:34-113, core/actions.py:5-15 ------------------------------
UNSEEN, EMPTY, WALL, FLOOR, DOOR, KEY, BALL, BOX, GOAL, LAVA, AGENT = range(11)
OPEN, CLOSED, LOCKED = 0, 1, 2
LEFT, RIGHT, FORWARD, PICKUP, DROP, TOGGLE, DONE = range(7)
DIR_TO_VEC = ((1, 0), (0, 1), (-1, 0), (0, -1))  # core/constants.py:21-30
WALL_ENCODING = (WALL, 5, 0)  # utils/obs.py:14  (grey wall)
A_DIR, A_X, A_Y, A_TERM, A_CT, A_CC, A_CS, A_COLOR = range(8)

HOOK_NONE, HOOK_BUP, HOOK_RBD, HOOK_LH = 0, 1, 2, 3
RED, BLUE = 0, 2  # core/constants.py:51-60

M64 = (1 << 64) - 1
M128 = (1 << 128) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645  # numpy PCG64 default multiplier


@dataclass
class OracleConfig:
    W: int
    H: int
    n: int
    V: int = 7
    max_steps: int = 100
    see_through_walls: bool = False      # agents[0]'s flag is used for all (base.py:364-365)
    allow_agent_overlap: bool = True     # base.py:95
    joint_reward: bool = False           # base.py:96
    success_any: bool = True             # success_termination_mode == 'any' (base.py:97)
    failure_any: bool = False            # failure_termination_mode == 'any' (base.py:98)
    hook: int = HOOK_NONE                # env-specific step() post-hook
    hook_param: int = 0                  # LockedHallway: number of rooms (= doors)
    auto_reset: bool = False             # engine extension ("next-step" reset), see DESIGN.md
    layout_stride: int = 1


# --- PCG64 (numpy Generator.random, call site base.py:399) ----------------------------------
def pcg64_next_double(state: int, inc: int) -> tuple[int, float]:
    """One `Generator(PCG64).random()` draw: LCG step, XSL-RR output, top 53 bits -> [0,1)."""
    state = (state * PCG_MULT + inc) & M128
    hi, lo = state >> 64, state & M64
    x = hi ^ lo
    rot = hi >> 58
    out = ((x >> rot) | (x << ((64 - rot) & 63))) & M64
    return state, (out >> 11) * (1.0 / 9007199254740992.0)


def _to_int128(words) -> int:
    return int(words[0]) | (int(words[1]) << 64)


def _from_int128(v: int, out) -> None:
    out[0] = np.uint64(v & M64)
    out[1] = np.uint64(v >> 64)


# --- observation: utils/obs.py ---------------------------------------------------------------
def gen_obs_env(cfg: OracleConfig, grid: np.ndarray, agents: np.ndarray) -> np.ndarray:
    """`gen_obs_grid_encoding` (utils/obs.py:66-102) for one env -> int8 (n, V, V, 3)."""
    n, V, W, H = cfg.n, cfg.V, cfg.W, cfg.H
    g = grid
    if n > 1:  # utils/obs.py:163-171: stamp non-terminated agents, ascending index
        g = grid.copy()
        for j in range(n):
            if not agents[j, A_TERM]:
                g[agents[j, A_X], agents[j, A_Y]] = (AGENT, agents[j, A_COLOR], agents[j, A_DIR])
    obs = np.zeros((n, V, V, 3), dtype=np.int8)
    half = V // 2
    for k in range(n):
        d = int(agents[k, A_DIR])
        px, py = int(agents[k, A_X]), int(agents[k, A_Y])
        # get_view_exts (utils/obs.py:276-316) + rotation (utils/obs.py:184-202)
        if d == 0:
            tx, ty = px, py - half
        elif d == 1:
            tx, ty = px - half, py
        elif d == 2:
            tx, ty = px - V + 1, py - half
        elif d == 3:
            tx, ty = px - half, py - V + 1
        else:
            tx, ty = 0, 0
        rot = (d + 1) % 4
        for i in range(V):
            for j in range(V):
                x, y = tx + i, ty + j
                if rot == 0:
                    ir, jr = i, j
                elif rot == 1:
                    ir, jr = j, V - i - 1
                elif rot == 2:
                    ir, jr = V - i - 1, V - j - 1
                else:
                    ir, jr = V - j - 1, i
                if 0 <= x < W and 0 <= y < H:
                    obs[k, ir, jr] = g[x, y]
                else:
                    obs[k, ir, jr] = WALL_ENCODING
        obs[k, half, V - 1] = agents[k, A_CT:A_CS + 1]  # utils/obs.py:207 (terminated or not)

    if cfg.see_through_walls:  # utils/obs.py:95
        return obs
    for k in range(n):
        # see_behind (utils/obs.py:47-63)
        t, s = obs[k, :, :, 0], obs[k, :, :, 2]
        see = ~((t == WALL) | ((t == DOOR) & (s != OPEN)))
        vis = np.zeros((V, V), dtype=bool)
        vis[half, V - 1] = True
        for j in range(V - 1, -1, -1):  # get_vis_mask (utils/obs.py:236-273)
            for i in range(0, V - 1):
                if vis[i, j] and see[i, j]:
                    vis[i + 1, j] = True
                    if j > 0:
                        vis[i + 1, j - 1] = True
                        vis[i, j - 1] = True
            for i in range(V - 1, 0, -1):
                if vis[i, j] and see[i, j]:
                    vis[i - 1, j] = True
                    if j > 0:
                        vis[i - 1, j - 1] = True
                        vis[i, j - 1] = True
        obs[k][~vis] = 0  # UNSEEN_ENCODING (utils/obs.py:15, 96-100)
    return obs


# --- transition: base.py:303-532 -------------------------------------------------------------
def _reward(step_count: int, max_steps: int) -> float:
    return 1 - 0.9 * (step_count / max_steps)  # base.py:598-602, float64


def _on_success(cfg, agents, k, rewards, terminated_out, step_count):  # base.py:478-507
    if cfg.success_any:
        agents[:, A_TERM] = 1
        terminated_out[:] = 1
    else:
        agents[k, A_TERM] = 1
        terminated_out[k] = 1
    if cfg.joint_reward:
        rewards[:] = _reward(step_count, cfg.max_steps)
    else:
        rewards[k] = _reward(step_count, cfg.max_steps)


def _on_failure(cfg, agents, k, terminated_out):  # base.py:509-532
    if cfg.failure_any:
        agents[:, A_TERM] = 1
        terminated_out[:] = 1
    else:
        agents[k, A_TERM] = 1
        terminated_out[k] = 1


def handle_actions_env(cfg, grid, agents, step_count, pcg_state, pcg_inc, actions, rewards,
                       cell_flags=None):
    """`MultiGridEnv.handle_actions` (base.py:378-476) for one env. Mutates grid/agents/pcg.

    `cell_flags[x, y] & 1`: the Door OBJECT at (x, y) is closed although `grid.state` shows it
    open. The reference keeps objects and the int array in sync with `grid.update()`, except in
    RedBlueDoorsEnv.step, which closes the blue door's object without updating the array
    (envs/redbluedoors.py:185): rules read the object, observations read the array."""
    n = cfg.n
    scratch_term = np.zeros(n, dtype=np.uint8)
    if n == 1:
        order = (0,)  # base.py:396-397
    else:  # base.py:399: np_random.random(size=n).argsort()
        s, inc = _to_int128(pcg_state), _to_int128(pcg_inc)
        draws = []
        for _ in range(n):
            s, u = pcg64_next_double(s, inc)
            draws.append(u)
        _from_int128(s, pcg_state)
        order = np.argsort(np.array(draws), kind="stable")
    for k in order:
        a = int(actions[k])
        if a < 0:  # agent id absent from the action dict (base.py:403-404)
            continue
        if agents[k, A_TERM]:  # base.py:408-409
            continue
        if a == LEFT:
            agents[k, A_DIR] = (int(agents[k, A_DIR]) - 1) % 4
            continue
        if a == RIGHT:
            agents[k, A_DIR] = (int(agents[k, A_DIR]) + 1) % 4
            continue
        if a == DONE:
            continue
        if a > DONE:
            raise ValueError(f"Unknown action: {a}")  # base.py:473-474
        dx, dy = DIR_TO_VEC[int(agents[k, A_DIR])]
        fx, fy = int(agents[k, A_X]) + dx, int(agents[k, A_Y]) + dy
        if not (0 <= fx < cfg.W and 0 <= fy < cfg.H):
            continue  # never reached in registered envs (outer wall ring); engine treats as no-op
        t, c, s_ = (int(v) for v in grid[fx, fy])
        if cell_flags is not None and t == DOOR and cell_flags[fx, fy] & 1:
            s_ = CLOSED  # what the Door object says
        agent_at_f = bool(((agents[:, A_X] == fx) & (agents[:, A_Y] == fy)).any())
        if a == FORWARD:  # base.py:420-436
            can_overlap = t in (EMPTY, FLOOR, GOAL, LAVA) or (t == DOOR and s_ == OPEN)
            if can_overlap:
                if not cfg.allow_agent_overlap and agent_at_f:
                    continue
                agents[k, A_X], agents[k, A_Y] = fx, fy
                if t == GOAL:
                    _on_success(cfg, agents, k, rewards, scratch_term, step_count)
                if t == LAVA:
                    _on_failure(cfg, agents, k, scratch_term)
        elif a == PICKUP:  # base.py:439-446
            if t in (KEY, BALL, BOX) and agents[k, A_CT] == EMPTY:
                agents[k, A_CT:A_CS + 1] = (t, c, s_)
                grid[fx, fy] = (EMPTY, 0, 0)
        elif a == DROP:  # base.py:449-459
            if agents[k, A_CT] != EMPTY and t == EMPTY and not agent_at_f:
                grid[fx, fy] = agents[k, A_CT:A_CS + 1]
                agents[k, A_CT:A_CS + 1] = (EMPTY, 0, 0)
        elif a == TOGGLE:  # base.py:462-467
            if t == DOOR:  # Door.toggle (core/world_object.py:458-474)
                if s_ == LOCKED:
                    if agents[k, A_CT] == KEY and agents[k, A_CC] == c:
                        grid[fx, fy, 2] = OPEN
                elif s_ == OPEN:
                    grid[fx, fy, 2] = CLOSED
                else:
                    grid[fx, fy, 2] = OPEN
                if cell_flags is not None:
                    cell_flags[fx, fy] &= 0xFE  # Door.toggle ends with grid.update(): in sync again
            elif t == BOX:  # Box.toggle (core/world_object.py:599-605); contains is None
                grid[fx, fy] = (EMPTY, 0, 0)


def step_env(cfg, grid, agents, step_count, pcg_state, pcg_inc, actions, cell_flags=None,
             hook_state=None):
    """`MultiGridEnv.step` (base.py:303-346) + env post-hook for one env.

    Returns (obs, reward f64 (n,), terminated u8 (n,), truncated bool, new_step_count).
    """
    n = cfg.n
    step_count = int(step_count) + 1  # base.py:333
    rewards = np.zeros(n, dtype=np.float64)
    handle_actions_env(cfg, grid, agents, step_count, pcg_state, pcg_inc, actions, rewards, cell_flags)
    obs = gen_obs_env(cfg, grid, agents)  # base.py:337
    terminated = agents[:, A_TERM].astype(np.uint8).copy()  # base.py:338
    truncated = step_count >= cfg.max_steps  # base.py:339
    if cfg.hook == HOOK_BUP:  # envs/blockedunlockpickup.py:166-175 (the env's only box)
        for k in range(n):
            if agents[k, A_CT] == BOX:
                _on_success(cfg, agents, k, rewards, terminated, step_count)
    if cfg.hook == HOOK_RBD:  # envs/redbluedoors.py:170-187
        for k in range(n):
            if actions[k] != TOGGLE:  # ids absent from the dict are -1
                continue
            dx, dy = DIR_TO_VEC[int(agents[k, A_DIR]) & 3]
            fx, fy = int(agents[k, A_X]) + dx, int(agents[k, A_Y]) + dy
            if not (0 <= fx < cfg.W and 0 <= fy < cfg.H):
                continue
            t, c, s = (int(v) for v in grid[fx, fy])
            if cell_flags is not None and cell_flags[fx, fy] & 1:
                s = CLOSED  # the object's state
            if (t, c, s) != (DOOR, BLUE, OPEN):  # fwd_obj == self.blue_door and it is open
                continue
            red = [(x, y) for x in range(cfg.W) for y in range(cfg.H)
                   if grid[x, y, 0] == DOOR and grid[x, y, 1] == RED]
            if red and grid[red[0][0], red[0][1], 2] == OPEN:
                _on_success(cfg, agents, k, rewards, terminated, step_count)
            else:
                _on_failure(cfg, agents, k, terminated)
                cell_flags[fx, fy] |= 1  # self.blue_door.is_open = False, WITHOUT grid.update()
    if cfg.hook == HOOK_LH:  # envs/locked_hallway.py:203-227
        # hook_state[0] = bit per door COLOUR already unlocked (doors have distinct colours for
        # num_rooms <= 6, envs/locked_hallway.py:156-158): the reference's `unlocked_doors` list
        for k in range(n):
            if actions[k] != TOGGLE:
                continue
            dx, dy = DIR_TO_VEC[int(agents[k, A_DIR]) & 3]
            fx, fy = int(agents[k, A_X]) + dx, int(agents[k, A_Y]) + dy
            if not (0 <= fx < cfg.W and 0 <= fy < cfg.H):
                continue
            t, c, s = (int(v) for v in grid[fx, fy])
            if t != DOOR or s == LOCKED:
                continue
            if not (int(hook_state[0]) >> c) & 1:
                hook_state[0] = int(hook_state[0]) | (1 << c)
                r = _reward(step_count, cfg.max_steps)
                if cfg.joint_reward:
                    rewards[:] = rewards + r  # rewards[k] += self._reward() for every k
                else:
                    rewards[k] = rewards[k] + r
        if bin(int(hook_state[0])).count("1") == cfg.hook_param:  # all doors unlocked
            terminated[:] = 1  # the returned dict only: agent state is NOT terminated
    return obs, rewards, terminated, truncated, step_count


def full_obs(grid, agents):
    """FullyObsWrapper.observation (multigrid/wrappers.py:50-58) for one env: Grid.encode() with every
    agent's (agent, colour, dir) on its cell, ascending agent index. `agents` = packed (n, 8)."""
    img = np.array(grid, dtype=np.int8)
    for k in range(agents.shape[0]):
        img[agents[k, A_X], agents[k, A_Y]] = (AGENT, agents[k, A_COLOR], agents[k, A_DIR])
    return img


def one_hot(x, dim_sizes=(11, 6, 4)):
    """OneHotObsWrapper.one_hot (multigrid/wrappers.py:158-190): (..., h, w, 3) ints -> uint8
    (..., h, w, sum(dim_sizes)); channel = offset of the dimension + value."""
    x = np.asarray(x)
    out = np.zeros(x.shape[:-1] + (sum(dim_sizes),), dtype=np.uint8)
    offset = 0
    for d, size in enumerate(dim_sizes):
        np.put_along_axis(out, (offset + x[..., d:d + 1]).astype(np.int64), 1, axis=-1)
        offset += size
    return out


# --- batched driver (mirrors the engine's fused mg_step_obs incl. its auto-reset extension) ---
def pack_agents(ref_agents: np.ndarray) -> np.ndarray:
    """Reference AgentState (..., 9) -> packed (..., 8)."""
    ref_agents = np.asarray(ref_agents)
    out = np.zeros(ref_agents.shape[:-1] + (8,), dtype=np.int8)
    out[..., A_DIR] = ref_agents[..., 2]
    out[..., A_X] = ref_agents[..., 3]
    out[..., A_Y] = ref_agents[..., 4]
    out[..., A_TERM] = ref_agents[..., 5]
    out[..., A_CT:A_CS + 1] = ref_agents[..., 6:9]
    out[..., A_COLOR] = ref_agents[..., 1]
    return out


def unpack_agents(packed: np.ndarray) -> np.ndarray:
    """Packed (..., 8) -> reference AgentState layout (..., 9)."""
    packed = np.asarray(packed)
    out = np.zeros(packed.shape[:-1] + (9,), dtype=np.int8)
    out[..., 0] = AGENT
    out[..., 1] = packed[..., A_COLOR]
    out[..., 2] = packed[..., A_DIR]
    out[..., 3] = packed[..., A_X]
    out[..., 4] = packed[..., A_Y]
    out[..., 5] = packed[..., A_TERM]
    out[..., 6:9] = packed[..., A_CT:A_CS + 1]
    return out


class OracleBatch:
    """B independent envs stepped one by one on the CPU (the checker for the CUDA path)."""

    def __init__(self, cfg: OracleConfig, grid, agents, pcg_state, pcg_inc,
                 pool_grid=None, pool_agents=None, layout_idx=None, step_count=None):
        self.cfg = cfg
        self.grid = np.array(grid, dtype=np.int8)
        self.agents = np.array(agents, dtype=np.int8)
        self.B = self.grid.shape[0]
        self.pcg_state = np.array(pcg_state, dtype=np.uint64)
        self.pcg_inc = np.array(pcg_inc, dtype=np.uint64)
        self.step_count = (np.zeros(self.B, np.int32) if step_count is None
                           else np.array(step_count, dtype=np.int32))
        self.pool_grid = None if pool_grid is None else np.array(pool_grid, dtype=np.int8)
        self.pool_agents = None if pool_agents is None else np.array(pool_agents, dtype=np.int8)
        self.layout_idx = (np.zeros(self.B, np.int32) if layout_idx is None
                           else np.array(layout_idx, dtype=np.int32))
        self.done = np.zeros(self.B, dtype=bool)
        self.cell_flags = np.zeros(self.grid.shape[:3], dtype=np.uint8)  # see handle_actions_env
        self.hook_state = np.zeros((self.B, 1), dtype=np.int32)

    def gen_obs(self):
        return np.stack([gen_obs_env(self.cfg, self.grid[b], self.agents[b])
                         for b in range(self.B)])

    def step(self, actions):
        cfg = self.cfg
        obs = np.zeros((self.B, cfg.n, cfg.V, cfg.V, 3), np.int8)
        rew = np.zeros((self.B, cfg.n), np.float64)
        term = np.zeros((self.B, cfg.n), np.uint8)
        trunc = np.zeros((self.B,), np.uint8)
        for b in range(self.B):
            done = bool(self.agents[b, :, A_TERM].all()) or self.step_count[b] >= cfg.max_steps
            if cfg.hook == HOOK_LH:  # its termination lives in the returned dict only (all doors unlocked)
                done = done or bin(int(self.hook_state[b, 0])).count("1") == cfg.hook_param
            if cfg.auto_reset and done:
                K = self.pool_grid.shape[0]
                self.layout_idx[b] = (int(self.layout_idx[b]) + cfg.layout_stride) % K
                self.grid[b] = self.pool_grid[self.layout_idx[b]]
                self.agents[b] = self.pool_agents[self.layout_idx[b]]
                self.step_count[b] = 0
                self.cell_flags[b] = 0
                self.hook_state[b] = 0
                obs[b] = gen_obs_env(cfg, self.grid[b], self.agents[b])
                continue
            o, r, t, tr, sc = step_env(cfg, self.grid[b], self.agents[b], self.step_count[b],
                                       self.pcg_state[b], self.pcg_inc[b], actions[b],
                                       self.cell_flags[b], self.hook_state[b])
            obs[b], rew[b], term[b], trunc[b] = o, r, t, tr
            self.step_count[b] = sc
        return obs, rew, term, trunc
