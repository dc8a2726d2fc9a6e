"""Host-side episode layout generation (the reset side of the hot path; stays in Python like the
reference's `_gen_grid`).

Each `*Layout` class is the batched engine's counterpart of one reference env class's
`_gen_grid` (multigrid/envs/*.py) and produces, per episode, the packed arrays the CUDA engine
consumes: `grid` int8 (W,H,3) and `agents` int8 (n,8). Layouts are generated on the host into a
*layout pool* that lives in HBM; the kernels copy from that pool on (auto-)reset.

Random draws follow the reference's order and generator split so that, given the same two
numpy generators, the layout is identical to the reference's (tests/test_layouts.py checks this
against states recorded from the reference):
  * `layout_rng`  = the generator RandomMixin captured at construction (utils/random.py:14-21,
    base.py:143): every `_rand_*` / place_obj / place_agent draw;
  * `order_rng`   = `env.np_random` (gymnasium): only RoomGrid door positions draw from it at
    reset time (core/roomgrid.py:324, 106-124) -- and the per-step agent order (base.py:399).
"""
from __future__ import annotations

from collections import deque

import numpy as np

from .core.constants import Color, Direction, State, Type

# packed agent record, see include/multigrid_b200.h
A_DIR, A_X, A_Y, A_TERM, A_CT, A_CC, A_CS, A_COLOR = range(8)

EMPTY_CELL = (Type.empty.to_index(), 0, 0)
WALL_CELL = (Type.wall.to_index(), Color.grey.to_index(), 0)
_DIRS = list(Direction)
_COLORS = list(Color)
_VEC = ((1, 0), (0, 1), (-1, 0), (0, -1))


def encode(kind: Type, color: Color = Color.red, state: State = State.open) -> tuple[int, int, int]:
    """(type, color, state) cell encoding (WorldObj.encode, core/world_object.py:235-247)."""
    return (Type(kind).to_index(), Color(color).to_index(), State(state).to_index())


class PlacementError(RecursionError):
    """Rejection sampling gave up (reference raises RecursionError, base.py:640-641)."""


class Canvas:
    """One episode's grid + agent records under construction, plus the two generators."""

    def __init__(self, width: int, height: int, num_agents: int,
                 layout_rng: np.random.Generator, order_rng: np.random.Generator):
        self.width, self.height, self.num_agents = width, height, num_agents
        self.rng, self.order_rng = layout_rng, order_rng
        self.grid = np.zeros((width, height, 3), dtype=np.int8)
        self.grid[...] = EMPTY_CELL  # Grid.__init__, core/grid.py:53-55
        self.agents = np.zeros((num_agents, 8), dtype=np.int8)
        self.agents[:, A_DIR] = -1           # AgentState defaults, core/agent.py:241-247
        self.agents[:, A_X:A_Y + 1] = -1
        self.agents[:, A_CT] = Type.empty.to_index()
        self.agents[:, A_COLOR] = np.arange(num_agents) % len(_COLORS)

    # -- RandomMixin equivalents (utils/random.py:23-103), same draw per call ------------------
    def rand_int(self, low: int, high: int) -> int:
        return int(self.rng.integers(low, high))

    def rand_bool(self) -> bool:
        return int(self.rng.integers(0, 2)) == 0

    def rand_elem(self, items):
        items = list(items)
        return items[self.rand_int(0, len(items))]

    def rand_color(self) -> Color:
        return self.rand_elem(_COLORS)

    def rand_perm(self, items) -> list:
        items = list(items)  # RandomMixin._rand_perm (utils/random.py:77-85): Generator.shuffle on a list
        self.rng.shuffle(items)
        return items

    # -- Grid drawing (core/grid.py:133-195) --------------------------------------------------
    def set(self, x: int, y: int, cell) -> None:
        self.grid[x, y] = EMPTY_CELL if cell is None else cell

    def is_empty(self, x: int, y: int) -> bool:
        return self.grid[x, y, 0] == Type.empty.to_index()

    def wall_rect(self, x: int, y: int, w: int, h: int) -> None:
        self.grid[x:x + w, y] = WALL_CELL
        self.grid[x:x + w, y + h - 1] = WALL_CELL
        self.grid[x, y:y + h] = WALL_CELL
        self.grid[x + w - 1, y:y + h] = WALL_CELL

    # -- placement (base.py:604-697) ----------------------------------------------------------
    def place_obj(self, cell, top=None, size=None, reject_fn=None, max_tries=float("inf")):
        top = (0, 0) if top is None else (max(top[0], 0), max(top[1], 0))
        size = (self.width, self.height) if size is None else size
        tries = 0
        while True:
            if tries > max_tries:
                raise PlacementError("rejection sampling failed in place_obj")
            tries += 1
            pos = (self.rand_int(top[0], min(top[0] + size[0], self.width)),
                   self.rand_int(top[1], min(top[1] + size[1], self.height)))
            if not self.is_empty(*pos):
                continue
            if ((self.agents[:, A_X] == pos[0]) & (self.agents[:, A_Y] == pos[1])).any():
                continue
            if reject_fn is not None and reject_fn(self, pos):
                continue
            break
        self.set(pos[0], pos[1], cell)
        return pos

    def place_agent(self, k: int, top=None, size=None, rand_dir=True, max_tries=float("inf")):
        self.agents[k, A_X:A_Y + 1] = -1
        pos = self.place_obj(None, top, size, max_tries=max_tries)
        self.agents[k, A_X:A_Y + 1] = pos
        if rand_dir:
            self.agents[k, A_DIR] = self.rand_int(0, 4)
        return pos

    def front_cell(self, k: int):
        dx, dy = _VEC[int(self.agents[k, A_DIR])]
        return self.grid[int(self.agents[k, A_X]) + dx, int(self.agents[k, A_Y]) + dy]

    def check(self) -> None:
        """The asserts of MultiGridEnv.reset (base.py:283-289)."""
        assert (self.agents[:, A_X:A_Y + 1] >= 0).all() and (self.agents[:, A_DIR] >= 0).all()
        for k in range(self.num_agents):
            t, _, s = self.grid[self.agents[k, A_X], self.agents[k, A_Y]]
            ok = t in (Type.empty.to_index(), Type.floor.to_index(), Type.goal.to_index(),
                       Type.lava.to_index()) or (t == Type.door.to_index() and s == 0)
            assert ok, "agent starts on a non-overlappable cell"


def _next_to_agents(canvas: Canvas, pos) -> bool:
    """reject_next_to (core/roomgrid.py:46-51): within distance 1 of any agent's position."""
    d = canvas.agents[:, A_X:A_Y + 1].astype(np.int64) - np.asarray(pos, dtype=np.int64)
    return bool(((d * d).sum(-1) <= 1).any())


class _Room:
    def __init__(self, top, size):
        self.top, self.size = top, size
        self.doors = {d: None for d in _DIRS}      # None | True (wall removed) | dict(door)
        self.neighbors = {d: None for d in _DIRS}
        self.objs = []

    @property
    def locked(self) -> bool:
        return any(isinstance(d, dict) and d["locked"] for d in self.doors.values())

    def door_pos(self, direction, random):
        """Room.set_door_pos (core/roomgrid.py:87-124); `random` is the ORDER generator."""
        left, top = self.top
        right, bottom = left + self.size[0] - 1, top + self.size[1] - 1
        if direction == Direction.right:
            return (right, int(random.integers(top + 1, bottom)) if random else (top + bottom) // 2)
        if direction == Direction.down:
            return (int(random.integers(left + 1, right)) if random else (left + right) // 2, bottom)
        if direction == Direction.left:
            return (left, int(random.integers(top + 1, bottom)) if random else (top + bottom) // 2)
        return (int(random.integers(left + 1, right)) if random else (left + right) // 2, top)


class RoomCanvas(Canvas):
    """Canvas with the room lattice of RoomGrid (core/roomgrid.py:139-495)."""

    def __init__(self, room_size, num_rows, num_cols, num_agents, layout_rng, order_rng):
        self.room_size, self.num_rows, self.num_cols = room_size, num_rows, num_cols
        super().__init__((room_size - 1) * num_cols + 1, (room_size - 1) * num_rows + 1,
                         num_agents, layout_rng, order_rng)
        step = room_size - 1
        self.rooms = [[_Room((c * step, r * step), (room_size, room_size))
                       for c in range(num_cols)] for r in range(num_rows)]
        for r in range(num_rows):
            for c in range(num_cols):
                room = self.rooms[r][c]
                self.wall_rect(*room.top, *room.size)
                if c < num_cols - 1:
                    room.neighbors[Direction.right] = self.rooms[r][c + 1]
                if r < num_rows - 1:
                    room.neighbors[Direction.down] = self.rooms[r + 1][c]
                if c > 0:
                    room.neighbors[Direction.left] = self.rooms[r][c - 1]
                if r > 0:
                    room.neighbors[Direction.up] = self.rooms[r - 1][c]
        # agents start in the middle room, facing right (core/roomgrid.py:231-236)
        self.agents[:, A_DIR] = Direction.right
        self.agents[:, A_X] = (num_cols // 2) * step + room_size // 2
        self.agents[:, A_Y] = (num_rows // 2) * step + room_size // 2

    def room(self, col, row) -> _Room:
        return self.rooms[row][col]

    def add_object(self, col, row, kind=None, color=None):
        kind = kind or self.rand_elem([Type.key, Type.ball, Type.box])
        color = color or self.rand_color()
        room = self.room(col, row)
        pos = self.place_obj(encode(kind, color), room.top, room.size,
                             reject_fn=_next_to_agents, max_tries=1000)
        room.objs.append((kind, color))
        return (kind, color), pos

    def add_door(self, col, row, direction=None, color=None, locked=None, rand_pos=True):
        room = self.room(col, row)
        if direction is None:
            while True:
                direction = self.rand_elem(_DIRS)
                if room.neighbors[direction] is not None and room.doors[direction] is None:
                    break
        else:
            assert room.neighbors[direction] is not None, "no neighbor in this direction"
            assert room.doors[direction] is None, "door already exists"
        color = color if color is not None else self.rand_color()
        locked = locked if locked is not None else self.rand_bool()
        pos = room.door_pos(direction, self.order_rng if rand_pos else None)
        self.set(pos[0], pos[1], encode(Type.door, color, State.locked if locked else State.closed))
        door = dict(color=color, locked=locked, pos=pos)
        room.doors[direction] = door
        room.neighbors[direction].doors[_DIRS[(direction + 2) % 4]] = door
        return door, pos

    def remove_wall(self, col, row, direction):
        room = self.room(col, row)
        assert room.doors[direction] is None and room.neighbors[direction]
        (tx, ty), (w, h) = room.top, room.size
        if direction == Direction.right:
            self.grid[tx + w - 1, ty + 1:ty + h - 1] = EMPTY_CELL
        elif direction == Direction.down:
            self.grid[tx + 1:tx + w - 1, ty + h - 1] = EMPTY_CELL
        elif direction == Direction.left:
            self.grid[tx, ty + 1:ty + h - 1] = EMPTY_CELL
        else:
            self.grid[tx + 1:tx + w - 1, ty] = EMPTY_CELL
        room.doors[direction] = True
        room.neighbors[direction].doors[_DIRS[(direction + 2) % 4]] = True

    def place_agent_in_room(self, k, col=None, row=None, rand_dir=True):
        col = col if col is not None else self.rand_int(0, self.num_cols)
        row = row if row is not None else self.rand_int(0, self.num_rows)
        room = self.room(col, row)
        while True:  # not right in front of an object (core/roomgrid.py:395-402)
            self.place_agent(k, room.top, room.size, rand_dir, max_tries=1000)
            t = self.front_cell(k)[0]
            if t == Type.empty.to_index() or t == Type.wall.to_index():
                break
        return tuple(self.agents[k, A_X:A_Y + 1])

    def connect_all(self, door_colors=_COLORS, max_itrs=5000):
        """Add unlocked doors until every room is reachable (core/roomgrid.py:406-452)."""
        total = self.num_rows * self.num_cols
        start = self.room(0, 0)
        for _ in range(max_itrs):
            seen, queue = set(), deque([start])
            while queue:
                room = queue.popleft()
                if id(room) in seen:
                    continue
                seen.add(id(room))
                queue.extend(room.neighbors[d] for d in _DIRS if room.doors[d] is not None)
            if len(seen) == total:
                return
            col, row = self.rand_int(0, self.num_cols), self.rand_int(0, self.num_rows)
            direction = self.rand_elem(_DIRS)
            room = self.room(col, row)
            other = room.neighbors[direction]
            if not other or room.doors[direction]:
                continue
            if room.locked or other.locked:
                continue
            self.add_door(col, row, direction=direction, color=self.rand_elem(door_colors),
                          locked=False)
        raise PlacementError("connect_all() failed")


# ---- per-env layout generators --------------------------------------------------------------------
class Layout:
    """Base: static description (size, max_steps, flags, hook, mission) + `generate()`."""
    hook = 0
    mission = "maximize reward"

    def __init__(self, width, height, num_agents, max_steps):
        self.width, self.height = width, height
        self.num_agents, self.max_steps = num_agents, max_steps

    #: True when every episode has the same layout and no RNG is consumed
    deterministic = False

    def generate(self, layout_rng, order_rng):
        """-> (grid int8 (W,H,3), agents int8 (n,8), info dict)"""
        raise NotImplementedError


class EmptyLayout(Layout):
    """EmptyEnv._gen_grid (envs/empty.py:151-170)."""
    mission = "get to the green goal square"

    def __init__(self, num_agents, size=8, agent_start_pos=(1, 1), agent_start_dir=Direction.right,
                 max_steps=None):
        super().__init__(size, size, num_agents, max_steps or 4 * size ** 2)
        self.agent_start_pos, self.agent_start_dir = agent_start_pos, agent_start_dir
        self.deterministic = agent_start_pos is not None and agent_start_dir is not None

    def generate(self, layout_rng, order_rng):
        c = Canvas(self.width, self.height, self.num_agents, layout_rng, order_rng)
        c.wall_rect(0, 0, self.width, self.height)
        c.set(self.width - 2, self.height - 2, encode(Type.goal, Color.green))
        for k in range(self.num_agents):
            if self.deterministic:
                c.agents[k, A_X:A_Y + 1] = self.agent_start_pos
                c.agents[k, A_DIR] = int(self.agent_start_dir)
            else:
                c.place_agent(k)
        c.check()
        return c.grid, c.agents, {}


class BlockedUnlockPickupLayout(Layout):
    """BlockedUnlockPickupEnv._gen_grid (envs/blockedunlockpickup.py:142-164)."""
    hook = 1  # MG_HOOK_BLOCKED_UNLOCK_PICKUP

    def __init__(self, num_agents, room_size=6, max_steps=None):
        assert room_size >= 4
        self.room_size = room_size
        super().__init__((room_size - 1) * 2 + 1, room_size, num_agents,
                         max_steps or 16 * room_size ** 2)

    def generate(self, layout_rng, order_rng):
        c = RoomCanvas(self.room_size, 1, 2, self.num_agents, layout_rng, order_rng)
        (kind, color), _ = c.add_object(1, 0, kind=Type.box)
        door, pos = c.add_door(0, 0, Direction.right, locked=True)
        c.set(pos[0] - 1, pos[1], encode(Type.ball, c.rand_color()))
        c.add_object(0, 0, Type.key, door["color"])
        for k in range(self.num_agents):
            c.place_agent_in_room(k, 0, 0)
        c.check()
        return c.grid, c.agents, dict(mission=f"pick up the {color.value} {kind.value}")


class PlaygroundLayout(Layout):
    """PlaygroundEnv._gen_grid (envs/playground.py:122-137)."""
    mission = ""

    def __init__(self, num_agents, room_size=7, num_rows=3, num_cols=3, max_steps=100):
        self.room_size, self.num_rows, self.num_cols = room_size, num_rows, num_cols
        super().__init__((room_size - 1) * num_cols + 1, (room_size - 1) * num_rows + 1,
                         num_agents, max_steps)

    def generate(self, layout_rng, order_rng):
        c = RoomCanvas(self.room_size, self.num_rows, self.num_cols, self.num_agents,
                       layout_rng, order_rng)
        c.connect_all()
        for _ in range(12):
            col = c.rand_int(0, self.num_cols)
            row = c.rand_int(0, self.num_rows)
            c.add_object(col, row)
        for k in range(self.num_agents):
            c.place_agent_in_room(k)
        c.check()
        return c.grid, c.agents, {}


class LockedHallwayLayout(Layout):
    """LockedHallwayEnv._gen_grid (envs/locked_hallway.py:150-194)."""
    hook = 3  # MG_HOOK_LOCKED_HALLWAY
    mission = "unlock all the doors"

    def __init__(self, num_agents, num_rooms=6, room_size=5, max_hallway_keys=1, max_keys_per_room=2,
                 max_steps=None):
        assert room_size >= 4 and num_rooms % 2 == 0
        assert num_rooms <= len(_COLORS), "door identity is tracked by colour: at most 6 rooms"
        self.num_rooms, self.room_size = num_rooms, room_size
        self.max_hallway_keys, self.max_keys_per_room = max_hallway_keys, max_keys_per_room
        self.hook_param = num_rooms
        super().__init__((room_size - 1) * 3 + 1, (room_size - 1) * (num_rooms // 2) + 1, num_agents,
                         max_steps or 8 * num_rooms * room_size ** 2)

    def generate(self, layout_rng, order_rng):
        LEFT, HALLWAY, RIGHT = range(3)
        num_rows = self.num_rooms // 2
        c = RoomCanvas(self.room_size, num_rows, 3, self.num_agents, layout_rng, order_rng)
        color_sequence = _COLORS * -(-self.num_rooms // len(_COLORS))
        color_sequence = c.rand_perm(color_sequence)[:self.num_rooms]
        for row in range(num_rows - 1):
            c.remove_wall(HALLWAY, row, Direction.down)
        rooms = {}
        door_colors = c.rand_perm(color_sequence)
        for row in range(num_rows):
            for col, direction in ((LEFT, Direction.right), (RIGHT, Direction.left)):
                color = door_colors.pop()
                rooms[color] = c.room(col, row)
                c.add_door(col, row, direction=direction, color=color, locked=True, rand_pos=False)
        num_hallway_keys = c.rand_int(1, self.max_hallway_keys + 1)
        hallway_top = c.room(HALLWAY, 0).top
        hallway_size = (c.room(HALLWAY, 0).size[0], c.height)
        for key_color in color_sequence[:num_hallway_keys]:
            c.place_obj(encode(Type.key, key_color), hallway_top, hallway_size)
        key_index = num_hallway_keys
        while key_index < len(color_sequence):
            room = rooms[color_sequence[key_index - 1]]
            num_room_keys = c.rand_int(1, self.max_keys_per_room + 1)
            for key_color in color_sequence[key_index:key_index + num_room_keys]:
                c.place_obj(encode(Type.key, key_color), room.top, room.size)
                key_index += 1
        for k in range(self.num_agents):  # MultiGridEnv.place_agent, not RoomGrid's (:193-194)
            c.place_agent(k, hallway_top, hallway_size)
        c.check()
        return c.grid, c.agents, {}


class RedBlueDoorsLayout(Layout):
    """RedBlueDoorsEnv._gen_grid (envs/redbluedoors.py:142-168)."""
    hook = 2  # MG_HOOK_RED_BLUE_DOORS
    mission = "open the red door then the blue door"

    def __init__(self, num_agents, size=8, max_steps=None):
        self.size = size
        super().__init__(2 * size, size, num_agents, max_steps or 20 * size ** 2)

    def generate(self, layout_rng, order_rng):
        W, H = self.width, self.height
        c = Canvas(W, H, self.num_agents, layout_rng, order_rng)
        room_top, room_size = (W // 4, 0), (W // 2, H)
        c.wall_rect(0, 0, W, H)
        c.wall_rect(*room_top, *room_size)
        for k in range(self.num_agents):
            c.place_agent(k, top=room_top, size=room_size)
        y = c.rand_int(1, H - 1)
        c.set(room_top[0], y, encode(Type.door, Color.red, State.closed))
        y = c.rand_int(1, H - 1)
        c.set(room_top[0] + room_size[0] - 1, y, encode(Type.door, Color.blue, State.closed))
        c.check()
        return c.grid, c.agents, {}


def generate_pool(layout: Layout, count: int, seed: int | None = None):
    """`count` episode layouts -> (grid (K,W,H,3), agents (K,n,8), infos). Deterministic layouts
    collapse to K=1."""
    if layout.deterministic:
        count = 1
    ss = np.random.SeedSequence(seed)
    grids, agents, infos = [], [], []
    for child in ss.spawn(count):
        a, b = child.spawn(2)
        g, ag, info = layout.generate(np.random.Generator(np.random.PCG64(a)),
                                      np.random.Generator(np.random.PCG64(b)))
        grids.append(g)
        agents.append(ag)
        infos.append(info)
    return np.stack(grids), np.stack(agents), infos
