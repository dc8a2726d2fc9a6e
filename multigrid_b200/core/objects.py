"""Object view of a grid cell (reference: multigrid/core/world_object.py:28-605). In the batched engine the int8
tensors ARE the state (the reference keeps `Grid.state` equal to its object cache on every
path the registered envs use); these classes are the reference's names for reading and writing single cells
(`env.grid.get(e, x, y)`, `env.grid.set(e, x, y, Door('red', is_locked=True))`) and for the rule predicates the
kernels hard-code (`can_overlap`, `can_pickup`, `can_contain`). An object is its 3-int encoding
(type, colour, state), like the reference's `WorldObj(np.ndarray)`."""
from __future__ import annotations

import numpy as np

from .constants import Color, State, Type


class WorldObj(np.ndarray):
    """int array [type, color, state] (world_object.py:28-137)."""
    TYPE, COLOR, STATE = 0, 1, 2
    dim = 3
    _TYPES: dict = {}

    def __new__(cls, type: str | None = None, color: str = Color.red):
        name = type if type is not None else cls.__name__.lower()
        obj = np.zeros(cls.dim, dtype=np.int64).view(cls)
        obj[WorldObj.TYPE] = Type(name).to_index()
        obj[WorldObj.COLOR] = Color(color).to_index()
        return obj

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        WorldObj._TYPES[Type(cls.__name__.lower()).to_index()] = cls

    def __bool__(self):
        return self.type != Type.empty

    def __repr__(self):
        return f"{self.__class__.__name__}(color={self.color})"

    def __eq__(self, other):  # identity, like the reference (world_object.py:128-129)
        return self is other

    def __hash__(self):
        return id(self)

    @staticmethod
    def empty() -> "WorldObj":
        return WorldObj(type=Type.empty)

    @staticmethod
    def from_array(arr):
        """(type, color, state) -> WorldObj instance, None for an empty cell (world_object.py:139-160)."""
        t = int(arr[WorldObj.TYPE])
        if t == Type.empty.to_index():
            return None
        if t not in WorldObj._TYPES:
            raise ValueError(f"Unknown object type: {t}")
        cls = WorldObj._TYPES[t]
        obj = np.zeros(cls.dim, dtype=np.int64).view(cls)
        obj[...] = np.asarray(arr, dtype=np.int64)[:3]
        return obj

    @property
    def type(self) -> Type:
        return Type.from_index(int(self[WorldObj.TYPE]))

    @property
    def color(self) -> Color:
        return Color.from_index(int(self[WorldObj.COLOR]))

    @color.setter
    def color(self, value):
        self[WorldObj.COLOR] = Color(value).to_index()

    @property
    def state(self) -> State:
        return State.from_index(int(self[WorldObj.STATE]))

    @state.setter
    def state(self, value):
        self[WorldObj.STATE] = State(value).to_index()

    def can_overlap(self) -> bool:  # world_object.py:197-201
        return self.type == Type.empty

    def can_pickup(self) -> bool:   # world_object.py:203-207
        return False

    def can_contain(self) -> bool:  # world_object.py:209-213
        return False

    def encode(self) -> tuple[int, int, int]:  # world_object.py:235-248
        return tuple(int(v) for v in self)

    @staticmethod
    def decode(type_idx: int, color_idx: int, state_idx: int):  # world_object.py:250-271
        return WorldObj.from_array((type_idx, color_idx, state_idx))


class Goal(WorldObj):
    def __new__(cls, color: str = Color.green):
        return super().__new__(cls, color=color)

    def can_overlap(self) -> bool:  # world_object.py:287
        return True


class Floor(WorldObj):
    def __new__(cls, color: str = Color.blue):
        return super().__new__(cls, color=color)

    def can_overlap(self) -> bool:  # world_object.py:314
        return True


class Lava(WorldObj):
    def __new__(cls):
        return super().__new__(cls, color=Color.red)

    def can_overlap(self) -> bool:  # world_object.py:339
        return True


class Wall(WorldObj):
    def __new__(cls, color: str = Color.grey):
        return super().__new__(cls, color=color)


class Door(WorldObj):
    """state: open 0 / closed 1 / locked 2 (world_object.py:392-474)."""

    def __new__(cls, color: str = Color.blue, is_open: bool = False, is_locked: bool = False):
        door = super().__new__(cls, color=color)
        door.is_open = is_open
        door.is_locked = is_locked
        return door

    @property
    def is_open(self) -> bool:
        return self.state == State.open

    @is_open.setter
    def is_open(self, value: bool):
        if value:
            self.state = State.open
        elif not self.is_locked:
            self.state = State.closed

    @property
    def is_locked(self) -> bool:
        return self.state == State.locked

    @is_locked.setter
    def is_locked(self, value: bool):
        if value:
            self.state = State.locked
        elif not self.is_open:
            self.state = State.closed

    def can_overlap(self) -> bool:  # world_object.py:452-456
        return self.is_open


class Key(WorldObj):
    def __new__(cls, color: str = Color.blue):
        return super().__new__(cls, color=color)

    def can_pickup(self) -> bool:  # world_object.py:518
        return True


class Ball(WorldObj):
    def __new__(cls, color: str = Color.blue):
        return super().__new__(cls, color=color)

    def can_pickup(self) -> bool:  # world_object.py:556
        return True


class Box(WorldObj):
    """`contains` is not representable in the 3-int encoding (every registered env uses None): a toggled box leaves
    an empty cell (world_object.py:599-605)."""

    def __new__(cls, color: str = Color.yellow, contains=None):
        if contains is not None:
            raise NotImplementedError("Box.contains is outside the tensor state (SURVEY.md section 8a A2)")
        return super().__new__(cls, color=color)

    def can_pickup(self) -> bool:   # world_object.py:587
        return True

    def can_contain(self) -> bool:
        return True
