"""Integer encodings shared by host and device (reference: multigrid/core/constants.py:21-113,
multigrid/core/actions.py:5-15, multigrid/utils/enum.py:42-89).

The member names, string values and index order are the reference's, because they ARE the wire
format: `grid.state[..., 0]` holds `Type` indices, `[..., 1]` `Color` indices, `[..., 2]`
`State` indices (or the direction for agent cells). The CUDA kernels hard-code the same numbers
(multigrid_b200/csrc/mg_kernels.cuh). Dynamic extension of the enums (`aenum.extend_enum`) is
out of scope: the device rule table is fixed at these 11 types.
"""
from __future__ import annotations

import enum

import numpy as np


class IndexedEnum(enum.Enum):
    """Enum whose members also have a dense integer index = definition order."""

    def to_index(self) -> int:
        return self._index_  # type: ignore[attr-defined]

    def __int__(self) -> int:
        return self.to_index()

    def __hash__(self):
        return hash(self.value)

    @classmethod
    def from_index(cls, index):
        """Member for an int, or an ndarray of member *values* for an array of ints."""
        members = list(cls)
        if np.ndim(index) == 0:
            return members[int(index)]
        values = np.array([m.value for m in members])
        return values[np.asarray(index)]

    @classmethod
    def _finalise(cls):
        for i, member in enumerate(cls):
            member._index_ = i
        return cls


class Type(str, IndexedEnum):
    unseen = "unseen"
    empty = "empty"
    wall = "wall"
    floor = "floor"
    door = "door"
    key = "key"
    ball = "ball"
    box = "box"
    goal = "goal"
    lava = "lava"
    agent = "agent"


class Color(str, IndexedEnum):
    red = "red"
    green = "green"
    blue = "blue"
    purple = "purple"
    yellow = "yellow"
    grey = "grey"

    @staticmethod
    def cycle(n: int) -> tuple["Color", ...]:
        members = list(Color)
        return tuple(members[i % len(members)] for i in range(int(n)))


class State(str, IndexedEnum):
    open = "open"
    closed = "closed"
    locked = "locked"


for _cls in (Type, Color, State):
    _cls._finalise()


class Direction(enum.IntEnum):
    right = 0
    down = 1
    left = 2
    up = 3

    def to_vec(self) -> np.ndarray:
        return DIR_TO_VEC[self]


#: (dx, dy) per direction: right = +x, down = +y
DIR_TO_VEC = [np.array((1, 0)), np.array((0, 1)), np.array((-1, 0)), np.array((0, -1))]


class Action(enum.IntEnum):
    left = 0      #: turn left
    right = 1     #: turn right
    forward = 2   #: move forward
    pickup = 3    #: pick up an object
    drop = 4      #: drop the carried object
    toggle = 5    #: toggle / activate an object
    done = 6      #: no-op


# minigrid-style lookup tables, as the reference exports them
OBJECT_TO_IDX = {t: t.to_index() for t in Type}
IDX_TO_OBJECT = {t.to_index(): t for t in Type}
COLOR_TO_IDX = {c: c.to_index() for c in Color}
IDX_TO_COLOR = {c.to_index(): c for c in Color}
STATE_TO_IDX = {s: s.to_index() for s in State}
COLOR_NAMES = sorted(list(Color))
