"""Device-resident batched state + the calls into the CUDA engine (C ABI, include/multigrid_b200.h).

`StepEngine` owns the HBM tensors of `num_envs` environments and advances them with the fused
sm_100a kernel. It is the array-level layer under `multigrid_b200.env.BatchedMultiGridEnv`
(the Gymnasium-style surface); both are host-side Python, as in the reference.

torch is used for device memory, streams and (optionally) torch.distributed only. There is no
CPU or PyTorch fallback for the computation: without the CUDA library and a CUDA device,
constructing a StepEngine raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _cabi


@dataclass
class EngineConfig:
    """Static per-batch configuration (mirrors MultiGridEnv.__init__ kwargs, base.py:85-207)."""
    width: int
    height: int
    num_agents: int
    view_size: int = 7
    max_steps: int = 100
    see_through_walls: bool = False
    allow_agent_overlap: bool = True
    joint_reward: bool = False
    success_termination_mode: str = "any"
    failure_termination_mode: str = "all"
    hook: int = _cabi.HOOK_NONE
    hook_param: int = 0  # LockedHallway: number of rooms
    auto_reset: bool = False
    layout_stride: int = 1
    stream_state: bool = False  # MG_FLAG_STREAM_STATE: L2 evict_first for state loads / obs stores (cache policy only)

    def __post_init__(self):
        if self.view_size % 2 != 1 or self.view_size < 3:  # core/agent.py:78-79
            raise AssertionError("agent_view_size must be odd and >= 3")
        if self.view_size > _cabi.MAX_VIEW:
            raise ValueError(f"agent_view_size > {_cabi.MAX_VIEW} is not supported")
        if not 1 <= self.num_agents <= _cabi.MAX_AGENTS:
            raise ValueError(f"num_agents must be in 1..{_cabi.MAX_AGENTS}")
        for mode in (self.success_termination_mode, self.failure_termination_mode):
            if mode not in ("any", "all"):
                raise ValueError(f"termination mode must be 'any' or 'all', got {mode!r}")

    @property
    def flags(self) -> int:
        return ((_cabi.FLAG_SEE_THROUGH_WALLS if self.see_through_walls else 0)
                | (_cabi.FLAG_ALLOW_OVERLAP if self.allow_agent_overlap else 0)
                | (_cabi.FLAG_JOINT_REWARD if self.joint_reward else 0)
                | (_cabi.FLAG_SUCCESS_ANY if self.success_termination_mode == "any" else 0)
                | (_cabi.FLAG_FAILURE_ANY if self.failure_termination_mode == "any" else 0)
                | (_cabi.FLAG_AUTO_RESET if self.auto_reset else 0)
                | (_cabi.FLAG_STREAM_STATE if self.stream_state else 0))


# Cell types a STATIC grid may hold (core/constants.py:34-48): empty, wall, floor, goal, lava. No action changes
# such a grid: pickup needs a key / ball / box, toggle a door / box, drop a carried object (base.py:439-467).
_STATIC_TYPES = (1, 2, 3, 8, 9)


def static_layout_ok(pool_grid, pool_agents) -> bool:
    """True when the promise of MG_FLAG_STATIC_GRID (include/multigrid_b200.h) holds for a batch whose envs all
    start from this pool: ONE layout of empty / wall / floor / goal / lava cells only, whose agents carry nothing
    and stand inside the grid on cells that are not walls. pool_grid (K,W,H,3), pool_agents (K,n,8) int8 arrays."""
    g, a = np.asarray(pool_grid), np.asarray(pool_agents)
    if g.shape[0] != 1 or not np.isin(g[..., 0], _STATIC_TYPES).all():
        return False
    W, H = g.shape[1:3]
    x, y = a[0, :, 1].astype(np.int64), a[0, :, 2].astype(np.int64)
    if ((x < 0) | (x >= W) | (y < 0) | (y >= H)).any() or (a[0, :, 4] != 1).any():
        return False
    if ((a[0, :, 0] < 0) | (a[0, :, 0] > 3)).any():
        return False
    return bool((g[0, x, y, 0] != 2).all())


def packed_obs_stride(view_size: int, bits: int = 9) -> int:
    """Bytes per agent of the packed wire formats (mg_packed_obs_stride / _bits): `bits` per cell, whole 8-byte words."""
    return ((bits * view_size * view_size + 63) // 64) * 8


def unpack_obs(packed, view_size: int, bits: int = 9, palette=None) -> np.ndarray:
    """Decoder of the packed observation wire formats (include/multigrid_b200.h, mg_pack_obs / mg_pack_obs_palette):
    uint8 [..., stride] -> int8 image [..., V, V, 3]. Cell (a, b) occupies bits [bits*(a*V + b), +bits) of the agent's
    little-endian record: with bits = 9 it IS the code type | colour << 4 | state << 7, with a palette (uint16 codes,
    `StepEngine.wire_palette()`) it is the index of that code. Host-side numpy; torch tensors are accepted."""
    if isinstance(packed, torch.Tensor):
        packed = packed.cpu().numpy()
    V, B = int(view_size), int(bits)
    p = np.ascontiguousarray(packed, dtype=np.uint8)
    assert p.shape[-1] == packed_obs_stride(V, B), (p.shape, V, B)
    words = p.view("<u8")                                   # [..., stride / 8]
    off = B * np.arange(V * V, dtype=np.int64)
    lo, sh = off // 64, (off % 64).astype(np.uint64)
    code = words[..., lo] >> sh
    spill = sh > np.uint64(64 - B)                           # the field continues in the next word
    hi = np.minimum(lo + 1, words.shape[-1] - 1)
    code = np.where(spill, code | (words[..., hi] << ((np.uint64(64) - sh) & np.uint64(63))), code) & np.uint64((1 << B) - 1)
    if palette is not None:
        code = np.asarray(palette, dtype=np.uint64)[code.astype(np.int64)]
    out = np.empty(code.shape + (3,), np.int8)
    out[..., 0] = code & np.uint64(15)
    out[..., 1] = (code >> np.uint64(4)) & np.uint64(7)
    out[..., 2] = code >> np.uint64(7)
    return out.reshape(code.shape[:-1] + (V, V, 3))


def wire_record_bytes(num_agents: int) -> int:
    """Bytes of one env record of the host wire (mg_wire_record_bytes)."""
    return (8 + 4 + 4 * ((num_agents + 7) // 8) + 7) & ~7


def unpack_wire(wire, num_envs: int, num_agents: int, view_size: int, bits: int, palette):
    """Decoder of the host wire (include/multigrid_b200.h, mg_step_obs_host_wire): one uint8 buffer ->
    (image int8 [E, n, V, V, 3], reward float64 [E, n], terminated bool [E, n], truncated bool [E]).
    reward[e, j] = the env's value added counts[e, j] times -- the additions are done here in float64, like the
    kernel that checked the record, so the rewards are the engine's bit for bit."""
    if isinstance(wire, torch.Tensor):
        wire = wire.cpu().numpy()
    E, n, V = int(num_envs), int(num_agents), int(view_size)
    w = np.ascontiguousarray(wire, dtype=np.uint8).reshape(-1)
    ps, rb = packed_obs_stride(V, bits), wire_record_bytes(n)
    obs_bytes = (E * n * ps + 15) & ~15
    image = unpack_obs(w[:E * n * ps].reshape(E, n, ps), V, bits, palette)
    rec = w[obs_bytes:obs_bytes + E * rb].reshape(E, rb)
    value = rec[:, :8].copy().view("<f8").reshape(E)
    tmask = rec[:, 8:12].copy().view("<u4").reshape(E)
    cw = rec[:, 12:12 + 4 * ((n + 7) // 8)].copy().view("<u4").reshape(E, -1)
    j = np.arange(n)
    counts = (cw[:, j >> 3] >> (4 * (j & 7)).astype(np.uint32)) & np.uint32(15)
    reward = np.zeros((E, n), np.float64)
    for k in range(1, int(counts.max(initial=0)) + 1):  # (1 almost always: repeated addition, not a multiply)
        reward = np.where(counts >= k, reward + value[:, None], reward)
    terminated = ((tmask[:, None] >> j.astype(np.uint32)) & np.uint32(1)).astype(bool)
    return image, reward, terminated, (tmask >> np.uint32(31)).astype(bool)


def _as_i64_bits(a) -> np.ndarray:
    """uint64 words -> int64 with the same bits (torch has no general uint64 support)."""
    return np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)


try:  # the raw current stream / device without building torch.cuda.Stream objects (~1.5 us per step saved)
    _raw_stream = torch._C._cuda_getCurrentRawStream
    _current_device = torch._C._cuda_getDevice
except AttributeError:  # pragma: no cover  (other torch builds)
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream
    _current_device = torch.cuda.current_device


class _Plan:
    """Owner of one MgStepPlan handle (destroyed with the object)."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle

    def __del__(self):
        try:
            if self.handle:
                self.lib.mg_step_plan_destroy(self.handle)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass
        self.handle = None


class StepEngine:
    """HBM-resident state of `num_envs` envs + fused step/observe launches.

    Layout (env-major): cells int32 (E,W+1,H+1) = cell words
    type|color<<8|state<<16|opaque<<31 with wall sentinels in the last row/column (`grid` is the
    zero-copy (E,W,H,3) int8 view of their low three bytes = Grid.state); agents int8 (E,n,8) =
    [dir,x,y,terminated,carry_type,carry_color,carry_state,color]; step_count int32 (E);
    pcg_state / pcg_inc int64-bits (E,2) [lo,hi]; layout_idx int32 (E).
    Outputs: obs int8 (E,n,stride) exposed as a (E,n,V,V,3) view; reward f64 (E,n);
    terminated uint8 (E,n); truncated uint8 (E).
    """

    def __init__(self, cfg: EngineConfig, num_envs: int, device: torch.device | str | int = "cuda",
                 pool_grid=None, pool_agents=None):
        self.lib = _cabi.load()  # raises if the CUDA library is not built
        if not torch.cuda.is_available():
            raise RuntimeError("multigrid_b200 needs a CUDA device (no CPU fallback)")
        self.cfg = cfg
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("multigrid_b200 state lives in GPU memory; device must be CUDA")
        E, n, V = self.num_envs, cfg.num_agents, cfg.view_size
        W, H = cfg.width, cfg.height
        self.obs_stride = _cabi.obs_agent_stride(V)
        dev = self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        self.cells = z((E, W + 1, H + 1), torch.int32)
        self.agents = z((E, n, 8), torch.int8)
        self.step_count = z((E,), torch.int32)
        self.pcg_state = z((E, 2), torch.int64)
        self.pcg_inc = z((E, 2), torch.int64)
        self.layout_idx = z((E,), torch.int32)
        self.hook_state = z((E,), torch.int32)
        self.actions = z((E, n), torch.int8)
        self.obs_buf = z((E, n, self.obs_stride), torch.int8)
        self.reward = z((E, n), torch.float64)
        self.terminated = z((E, n), torch.uint8)
        self.truncated = z((E,), torch.uint8)
        self.status = z((1,), torch.int32)
        # chain tickets (include/multigrid_b200.h, MG_FLAG_CHAINED): maintained by chained step launches
        # per env {next ticket, tickets done, grid dirty, -}: chain tickets (MG_FLAG_CHAINED) and the flag of the
        # single-layout dedup (1 = the env's grid may differ from its pool layout)
        self.chain = z((E, 4), torch.int32)
        self.chain[:, 2] = 1
        self.pool_rep = None
        self._fresh = None  # MgLayoutGen of enable_fresh_layouts()
        self.one_hot = None  # (E, n, V, V, 21) uint8 once enable_one_hot() asked the step kernel for it
        self._palette = None  # (bits, codes, device lut) of the palette wire format, see wire_palette()
        # static-grid path (MG_FLAG_STATIC_GRID): memoised per-(x, y, dir) views of the single pool layout, and
        # whether the batch is known to satisfy the promise (True / False; None = injected state, check lazily)
        self.static_obs = None
        self._static_state = False
        self.use_static = True  # set False to force the general kernel (comparisons, tests)
        self._chain_armed = None  # the stream whose last operation on this engine was a step launch
        self.pool_grid = None
        self.pool_agents = None
        self._pool_rng = None
        if pool_grid is not None:
            self.set_layout_pool(pool_grid, pool_agents)
        self._host = None
        self._c = None
        self._plans = {}    # prepared launches per variant (mg_step_plan_*), dropped whenever the structs are rebuilt
        self._plan_run = self.lib.mg_step_plan_run
        self._act_shape = torch.Size((self.num_envs, self.cfg.num_agents))
        self._views = None  # (obs, reward, terminated, truncated): views of fixed buffers, built once
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())

    # -- grid layout ---------------------------------------------------------------------------
    @property
    def grid(self) -> torch.Tensor:
        """(E, W, H, 3) int8 = Grid.state (core/grid.py:54): zero-copy view of bytes 0..2 of the
        cell words. Read freely; WRITE through `load_state(grid=...)`, which also maintains the
        opaque bit the observation kernel reads."""
        W, H = self.cfg.width, self.cfg.height
        return self.cells.view(torch.int8).view(self.num_envs, W + 1, H + 1, 4)[:, :W, :H, :3]

    def _pack(self, grid3, out_cells: torch.Tensor) -> None:
        """mg_pack_grid: (K,W,H,3) bytes (numpy / torch, host / device) -> cell words in `out_cells`."""
        self._chain_armed = None
        W, H = self.cfg.width, self.cfg.height
        if isinstance(grid3, torch.Tensor):
            g = grid3.to(self.device, torch.int8).contiguous()
        else:
            g = torch.as_tensor(np.ascontiguousarray(grid3, dtype=np.int8)).to(self.device)
        K = out_cells.shape[0]
        assert g.numel() == K * W * H * 3, (tuple(g.shape), K, W, H)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.mg_pack_grid(W, H, K, g.data_ptr(), out_cells.data_ptr(), self._stream()),
                        "mg_pack_grid")

    # -- configuration / state injection ------------------------------------------------------
    def set_layout_pool(self, pool_grid, pool_agents) -> None:
        """Reset layouts: grid (K,W,H,3) int8 and packed agents (K,n,8) int8."""
        self._chain_armed = None
        cfg = self.cfg
        pg = np.ascontiguousarray(pool_grid, dtype=np.int8)
        pa = torch.as_tensor(np.ascontiguousarray(pool_agents, dtype=np.int8))
        assert pg.shape[1:] == (cfg.width, cfg.height, 3), pg.shape
        assert pa.shape == (pg.shape[0], cfg.num_agents, 8), pa.shape
        self.pool_grid = torch.zeros((pg.shape[0], cfg.width + 1, cfg.height + 1), dtype=torch.int32,
                                     device=self.device)
        self._pack(pg, self.pool_grid)
        self.pool_agents = pa.to(self.device)
        self._pool_rng = None
        self._pool_changed()

    def gen_layout_pool_red_blue_doors(self, size, rng_state, rng_inc, rng_buf=None):
        """mg_gen_layouts_red_blue_doors (envs/redbluedoors.py:142-168); arguments and result as
        gen_layout_pool_empty_random. refresh_layout_pool() continues the same generators."""
        self._chain_armed = None
        assert (self.cfg.width, self.cfg.height) == (2 * size, size)
        return self.gen_layout_pool_empty_random(rng_state, rng_inc, rng_buf, _family=("rbd", size))

    def gen_layout_pool_locked_hallway(self, num_rooms, room_size, max_hallway_keys, max_keys_per_room,
                                       rng_state, rng_inc, rng_buf=None):
        """mg_gen_layouts_locked_hallway (envs/locked_hallway.py:150-194); result as
        gen_layout_pool_empty_random."""
        self._chain_armed = None
        return self.gen_layout_pool_empty_random(
            rng_state, rng_inc, rng_buf, _family=("lh", num_rooms, room_size, max_hallway_keys, max_keys_per_room))

    def gen_layout_pool_empty_random(self, rng_state, rng_inc, rng_buf=None, _family=("empty", 0)):
        """mg_gen_layouts_empty_random: fill the reset-layout pool ON THE DEVICE with K =
        len(rng_state) EmptyEnv layouts with random agent placement (envs/empty.py:151-170), one per
        numpy PCG64 generator given as uint64 words (state [K,2], inc [K,2], optional buffered-uint32
        word [K]). Returns the advanced (state, buf) as numpy uint64 arrays."""
        self._chain_armed = None
        cfg = self.cfg
        K = len(rng_state)
        dev = self.device
        st = torch.as_tensor(_as_i64_bits(rng_state)).to(dev).reshape(K, 2).contiguous()
        inc = torch.as_tensor(_as_i64_bits(rng_inc)).to(dev).reshape(K, 2).contiguous()
        buf = torch.as_tensor(_as_i64_bits(np.zeros(K, np.uint64) if rng_buf is None else rng_buf)).to(dev)
        cells = torch.empty((K, cfg.width + 1, cfg.height + 1), dtype=torch.int32, device=dev)
        agents = torch.empty((K, cfg.num_agents, 8), dtype=torch.int8, device=dev)
        self.pool_grid, self.pool_agents = cells, agents
        self._pool_rng = (st, inc, buf)  # stays on the device: refresh_layout_pool() continues these streams
        self._pool_gen = (st, inc, buf, None, None)
        self._pool_family = _family
        self._c = None
        self.refresh_layout_pool()
        if int(self.status.item()) & 2:
            self.status.zero_()
            raise RecursionError("rejection sampling failed in place_obj")  # base.py:640-641
        return st.cpu().numpy().view(np.uint64), buf.cpu().numpy().view(np.uint64)

    def gen_layout_pool_playground(self, room_size, num_rows, num_cols, rng_state, rng_inc, rng_buf, order_state,
                                   order_inc):
        """mg_gen_layouts_playground (envs/playground.py:122-137); arguments / result as gen_layout_pool_bup
        (the info array is all zero: the mission is constant)."""
        self._chain_armed = None
        return self.gen_layout_pool_bup(room_size, rng_state, rng_inc, rng_buf, order_state, order_inc,
                                        _grid=(num_rows, num_cols))

    def gen_layout_pool_bup(self, room_size, rng_state, rng_inc, rng_buf, order_state, order_inc, _grid=None):
        """mg_gen_layouts_bup: fill the pool on the device with K BlockedUnlockPickup layouts
        (envs/blockedunlockpickup.py:142-164). Layout generators as in gen_layout_pool_empty_random; the
        ORDER generators (env.np_random of the K envs, uint64 words state/inc [K,2]) give the door heights.
        Returns (order_state [K,2], box colour index [K] int32, rng_state [K,2], rng_buf [K]) after the draws."""
        self._chain_armed = None
        cfg = self.cfg
        K, dev = len(rng_state), self.device
        t64 = lambda a: torch.as_tensor(_as_i64_bits(a)).to(dev).contiguous()  # noqa: E731
        st, inc = t64(rng_state).reshape(K, 2), t64(rng_inc).reshape(K, 2)
        buf = t64(np.zeros(K, np.uint64) if rng_buf is None else rng_buf)
        ost, oinc = t64(order_state).reshape(K, 2), t64(order_inc).reshape(K, 2)
        obuf = torch.zeros((K,), dtype=torch.int64, device=dev)  # buffered 32-bit half of the order streams
        rows, cols = _grid or (1, 2)
        assert (cfg.width, cfg.height) == (cols * (room_size - 1) + 1, rows * (room_size - 1) + 1)
        cells = torch.empty((K, cfg.width + 1, cfg.height + 1), dtype=torch.int32, device=dev)
        agents = torch.empty((K, cfg.num_agents, 8), dtype=torch.int8, device=dev)
        info = torch.zeros((K,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            if _grid is not None:
                _cabi.check(self.lib.mg_gen_layouts_playground(
                    room_size, rows, cols, cfg.num_agents, K, st.data_ptr(), inc.data_ptr(), buf.data_ptr(),
                    ost.data_ptr(), oinc.data_ptr(), obuf.data_ptr(), cells.data_ptr(), agents.data_ptr(),
                    self.status.data_ptr(), self._stream()), "mg_gen_layouts_playground")
            else:
                _cabi.check(self.lib.mg_gen_layouts_bup(
                    room_size, cfg.num_agents, K, st.data_ptr(), inc.data_ptr(), buf.data_ptr(), ost.data_ptr(),
                    oinc.data_ptr(), obuf.data_ptr(), cells.data_ptr(), agents.data_ptr(), info.data_ptr(),
                    self.status.data_ptr(), self._stream()), "mg_gen_layouts_bup")
        if int(self.status.item()) & 2:
            self.status.zero_()
            raise RecursionError("rejection sampling failed in place_obj")  # base.py:640-641
        self.pool_grid, self.pool_agents = cells, agents
        self._pool_rng = None  # (refresh_layout_pool is for the single-generator families)
        self._pool_gen = (st, inc, buf, info, obuf)  # the generators stay on the device for fresh layouts on auto-reset
        self._pool_family = ("pg", room_size, rows, cols) if _grid is not None else ("bup", room_size)
        self._pool_changed()
        u64 = lambda t: t.cpu().numpy().view(np.uint64)  # noqa: E731
        return u64(ost), info.cpu().numpy(), u64(st), u64(buf)

    def refresh_layout_pool(self) -> None:
        """Overwrite the pool with the NEXT layout of every pool generator (one kernel launch, no host
        work, asynchronous): with auto_reset, envs that reset after this draw from fresh layouts
        instead of cycling the first K. Only after gen_layout_pool_empty_random."""
        self._chain_armed = None
        if getattr(self, "_pool_rng", None) is None:
            raise RuntimeError("the layout pool was not generated on the device")
        cfg = self.cfg
        st, inc, buf = self._pool_rng
        tail = (st.shape[0], st.data_ptr(), inc.data_ptr(), buf.data_ptr(), self.pool_grid.data_ptr(),
                self.pool_agents.data_ptr(), self.status.data_ptr(), self._stream())
        with torch.cuda.device(self.device):
            if self._pool_family[0] == "lh":
                _cabi.check(self.lib.mg_gen_layouts_locked_hallway(*self._pool_family[1:], cfg.num_agents, *tail),
                            "mg_gen_layouts_locked_hallway")
            elif self._pool_family[0] == "rbd":
                _cabi.check(self.lib.mg_gen_layouts_red_blue_doors(self._pool_family[1], cfg.num_agents, *tail),
                            "mg_gen_layouts_red_blue_doors")
            else:
                _cabi.check(self.lib.mg_gen_layouts_empty_random(cfg.width, cfg.height, cfg.num_agents, *tail),
                            "mg_gen_layouts_empty_random")
        self._pool_changed()

    def enable_fresh_layouts(self) -> None:
        """A NEW layout for every episode under auto_reset, like the reference's reset() (base.py:250-301): needs a
        device-generated pool with one slot per env (K == num_envs, generator e = env e's RandomMixin generator).
        From now on every step launch is followed by mg_refresh_done_layouts, which regenerates the slot of each env
        that is done (and will be reset by the next launch) from that env's generator."""
        gen = getattr(self, "_pool_gen", None)
        if gen is None or self.pool_grid is None or self.pool_grid.shape[0] != self.num_envs:
            raise RuntimeError("fresh layouts need a device-generated layout pool with one slot per env")
        if not self.cfg.auto_reset:
            raise RuntimeError("fresh layouts are a mode of auto_reset")
        fam = self._pool_family
        code = {"empty": _cabi.LAYOUT_EMPTY_RANDOM, "bup": _cabi.LAYOUT_BUP, "rbd": _cabi.LAYOUT_RED_BLUE_DOORS,
                "lh": _cabi.LAYOUT_LOCKED_HALLWAY, "pg": _cabi.LAYOUT_PLAYGROUND}[fam[0]]
        params = [int(v) for v in fam[1:]] if fam[0] != "empty" else []
        params += [0] * (4 - len(params))
        st, inc, buf, info, obuf = gen
        self._fresh = _cabi.MgLayoutGen(code, (C.c_int32 * 4)(*params), st.data_ptr(), inc.data_ptr(), buf.data_ptr(),
                                        None if obuf is None else obuf.data_ptr(),
                                        None if info is None else info.data_ptr())
        self.cfg.layout_stride = 0  # an env always resets from ITS slot
        self.layout_idx.copy_(torch.arange(self.num_envs, dtype=torch.int32, device=self.device))
        self._c = None

    def _refresh_done(self, stream) -> None:
        c, st, _ = self._c
        rc = self.lib.mg_refresh_done_layouts(C.byref(c), self.num_envs, C.byref(st), C.byref(self._fresh),
                                              self.status.data_ptr(), stream)
        if rc:
            _cabi.check(rc, "mg_refresh_done_layouts")

    @property
    def grid_dirty(self) -> torch.Tensor:
        """(E,) int32 view: 1 = the env's grid may differ from its pool layout (word 2 of the chain records)."""
        return self.chain[:, 2]

    def _pool_changed(self) -> None:
        """The layout pool was replaced: no env is known to equal its layout any more; with a single layout,
        (re)build the 32-copy buffer clean groups load their cells from."""
        self.grid_dirty.fill_(1)
        self._palette = None
        self.pool_rep = (self.pool_grid[:1].repeat(32, 1, 1).contiguous()
                         if self.pool_grid is not None and self.pool_grid.shape[0] == 1 else None)
        self._c = None
        self._build_static()

    def _build_static(self) -> None:
        """(Re)build the memoised observations of a single static layout (mg_build_static_obs); without one,
        the engine only ever takes the general kernel."""
        self.static_obs, self._static_state = None, False
        cfg = self.cfg
        if self.pool_grid is None or self.pool_grid.shape[0] != 1 or cfg.hook != _cabi.HOOK_NONE:
            return
        if os.environ.get("MG_NO_STATIC") == "1":  # knob (comparisons): never build the table, general kernel only
            return
        W, H = cfg.width, cfg.height
        cells = self.pool_grid.cpu().numpy().view(np.uint32)
        grid3 = np.stack([cells & 0xff, (cells >> 8) & 0xff, (cells >> 16) & 0xff], -1)[:, :W, :H]
        if not static_layout_ok(grid3, self.pool_agents.cpu().numpy()):
            return
        # observation slots of the table's entry stride (3*V*V rounded up to 16 bytes: 160 for V = 7): the static
        # kernel then moves whole 16-byte pieces; `obs` stays the (E, n, V, V, 3) view of the first 3*V*V bytes
        stride = self.lib.mg_static_obs_stride(_cabi.obs_agent_stride(cfg.view_size))
        if stride != self.obs_stride:
            self.obs_stride = stride
            self.obs_buf = torch.zeros((self.num_envs, cfg.num_agents, stride), dtype=torch.int8, device=self.device)
            self._views, self._host = None, None
        table = torch.zeros((self.lib.mg_static_obs_bytes(W, H, stride),), dtype=torch.int8, device=self.device)
        c = _cabi.MgConfig(W, H, cfg.num_agents, cfg.view_size, cfg.max_steps, cfg.flags, cfg.hook,
                           self.obs_stride, 1, cfg.layout_stride, cfg.hook_param)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.mg_build_static_obs(C.byref(c), self.pool_grid.data_ptr(), table.data_ptr(),
                                                     self._stream()), "mg_build_static_obs")
        self.static_obs = table
        self._static_state = None  # the layout qualifies; whether the STATE does is checked on first use

    def _static_ok(self) -> bool:
        """Does the batch satisfy MG_FLAG_STATIC_GRID right now? After a reset from the pool it does by
        construction and no step can break it; after state injection (load_state) it is checked once on the
        device: every grid equals the layout, every agent is inside the grid, off the walls, empty-handed."""
        if self.static_obs is None or not self.use_static:
            return False
        if self._static_state is None:
            W, H = self.cfg.width, self.cfg.height
            a = self.agents.to(torch.int64)
            x, y = a[..., 1], a[..., 2]
            inside = (x >= 0) & (x < W) & (y >= 0) & (y < H) & (a[..., 4] == 1) & (a[..., 0] >= 0) & (a[..., 0] <= 3)
            under = self.pool_grid[0].to(torch.int64)[x.clamp(0, W - 1), y.clamp(0, H - 1)] & 0xff
            ok = inside.all() & (under != 2).all() & (self.cells == self.pool_grid[0]).all()
            self._static_state = bool(ok.item())
        return self._static_state

    def load_state(self, grid=None, agents=None, step_count=None, pcg_state=None, pcg_inc=None,
                   layout_idx=None, hook_state=None) -> None:
        """Inject state (numpy or torch, host or device). uint64 PCG words are passed as numpy."""
        self._chain_armed = None
        def put(dst, src, bits64=False):
            if src is None:
                return
            if isinstance(src, torch.Tensor):
                dst.copy_(src.to(dst.dtype).reshape(dst.shape))
            else:
                arr = _as_i64_bits(src) if bits64 else np.ascontiguousarray(src)
                dst.copy_(torch.as_tensor(arr).to(dst.dtype).reshape(dst.shape))
        if grid is not None:
            self._pack(grid, self.cells)
            self.grid_dirty.fill_(1)
        if grid is not None or agents is not None:
            self._palette = None
        if (grid is not None or agents is not None) and self.static_obs is not None:
            self._static_state = None  # re-checked on the device before the next step
        put(self.agents, agents)
        put(self.step_count, step_count)
        put(self.pcg_state, pcg_state, bits64=True)
        put(self.pcg_inc, pcg_inc, bits64=True)
        put(self.layout_idx, layout_idx)
        put(self.hook_state, hook_state)

    def reset_from_pool(self, layout_idx=None) -> None:
        """Host-driven reset of every env from the layout pool (step_count := 0)."""
        self._chain_armed = None
        if layout_idx is not None:
            self.load_state(layout_idx=layout_idx)
        idx = self.layout_idx.long()
        self.cells.copy_(self.pool_grid[idx])
        self.agents.copy_(self.pool_agents[idx])
        self.step_count.zero_()
        self.hook_state.zero_()
        self.grid_dirty.zero_()
        if self.static_obs is not None:
            self._static_state = True  # every env holds the one layout and its agents

    def reset_where(self, mask: torch.Tensor) -> None:
        """mg_reset_where: envs with a non-zero mask entry ((E,) bool / uint8 on the device) take the next
        layout of the pool, like the kernel's auto-reset does. Asynchronous."""
        self._chain_armed = None
        if mask.shape != (self.num_envs,) or mask.device != self.device:
            raise TypeError("mask must be a (num_envs,) tensor on the engine's device")
        m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
        m = m.contiguous()
        c, st, _ = self._structs()
        if self.pool_grid is None:
            raise RuntimeError("reset_where needs a layout pool (set_layout_pool)")
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.mg_reset_where(C.byref(c), self.num_envs, C.byref(st), m.data_ptr(), self._stream()),
                        "mg_reset_where")

    # -- C structs ---------------------------------------------------------------------------
    def _structs(self):
        if self._c is None:
            self._plans = {}  # (they point into the structs built below)
            cfg = self.cfg
            if cfg.auto_reset and self.pool_grid is None:
                raise RuntimeError("auto_reset needs a layout pool (set_layout_pool)")
            K = 0 if self.pool_grid is None else int(self.pool_grid.shape[0])
            c = _cabi.MgConfig(cfg.width, cfg.height, cfg.num_agents, cfg.view_size,
                               cfg.max_steps, cfg.flags, cfg.hook, self.obs_stride, K,
                               cfg.layout_stride, cfg.hook_param)
            p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
            st = _cabi.MgState(p(self.cells), p(self.agents), p(self.step_count), p(self.pcg_state),
                               p(self.pcg_inc), p(self.layout_idx), p(self.pool_grid),
                               p(self.pool_agents), p(self.hook_state), p(self.pool_rep), p(self.chain),
                               p(self.static_obs))
            out = _cabi.MgStepOut(p(self.obs_buf), p(self.reward), p(self.terminated),
                                  p(self.truncated), p(self.status), p(self.one_hot))
            mk = lambda extra: _cabi.MgConfig(cfg.width, cfg.height, cfg.num_agents, cfg.view_size,  # noqa: E731
                                              cfg.max_steps, cfg.flags | extra, cfg.hook, self.obs_stride, K,
                                              cfg.layout_stride, cfg.hook_param)
            cc, ch = mk(_cabi.FLAG_CHAINED), mk(_cabi.FLAG_CHAINED | _cabi.FLAG_CHAIN_HEAD)
            cs = mk(_cabi.FLAG_STATIC_GRID)
            self._c = (c, st, out)
            self._refs = (C.byref(c), C.byref(st), C.byref(out))
            self._chained_cfg = (cc, C.byref(cc), ch, C.byref(ch))
            self._static_cfg = (cs, C.byref(cs))
        return self._c

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def enable_one_hot(self) -> torch.Tensor:
        """From now on every fused step launch ALSO writes OneHotObsWrapper.one_hot of its observations
        (wrappers.py:158-190) into `self.one_hot` (E, n, V, V, 21) uint8, straight from the kernel's shared-memory
        stage (MgStepOut.one_hot): no second pass over `obs`. gen_obs() fills it with the standalone kernel."""
        if self.one_hot is None:
            V = self.cfg.view_size
            self.one_hot = torch.zeros((self.num_envs, self.cfg.num_agents, V, V, 21), dtype=torch.uint8,
                                       device=self.device)
            self._c = None
        return self.one_hot

    @property
    def obs(self) -> torch.Tensor:
        """(E, n, V, V, 3) int8 view of the padded observation buffer (zero-copy)."""
        V = self.cfg.view_size
        return self.obs_buf[:, :, :3 * V * V].unflatten(2, (V, V, 3))

    @property
    def direction(self) -> torch.Tensor:
        """(E, n) int8: the 'direction' observation aliases the agent-state tensor."""
        return self.agents[:, :, 0]

    # -- launches ------------------------------------------------------------------------------
    def gen_obs(self) -> torch.Tensor:
        """mg_gen_obs: observations of the current state (used after reset)."""
        self._chain_armed = None
        c, st, out = self._structs()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.mg_gen_obs(C.byref(c), self.num_envs, self.cells.data_ptr(),
                                            self.agents.data_ptr(), self.obs_buf.data_ptr(),
                                            self._stream()), "mg_gen_obs")
            if self.one_hot is not None:  # (reset-time observations: the standalone pass)
                _cabi.check(self.lib.mg_one_hot(self.cfg.view_size, self.num_envs * self.cfg.num_agents,
                                                self.obs_stride, self.obs_buf.data_ptr(), self.one_hot.data_ptr(),
                                                self._stream()), "mg_one_hot")
        return self.obs

    def step(self, actions: torch.Tensor | None = None, fused: bool = True, chained: bool = False):
        """mg_step_obs on device-resident int8 actions (E,n); -1 = agent absent.

        Returns device views (obs, reward, terminated, truncated); they are overwritten by the
        next call. Asynchronous on the current CUDA stream.

        chained=True (MG_FLAG_CHAINED, scheduling only): this launch is ordered after the previous step
        launch of this engine env by env instead of waiting for that whole grid. The caller promises that
        since that launch nothing else on the stream wrote `actions` or touched this engine's state or
        outputs (open-loop action tapes). The first chained launch after any other operation of the engine
        (plain step, load_state, gen_obs, reset_where, ...) is the head of a new chain and waits like a
        plain launch; plain launches never touch the tickets. It pays when launches on DIFFERENT engines are
        interleaved on the stream (several env batches in flight); on one engine stepped back to back every
        env of launch k+1 waits for launch k anyway, and chained stepping is no faster or slower than plain.
        """
        if actions is None:
            actions = self.actions
        if (actions.dtype != torch.int8 or actions.shape != self._act_shape or actions.device != self.device
                or not actions.is_contiguous()):
            raise TypeError("actions must be a contiguous int8 CUDA tensor of shape (num_envs, n)")
        if self._c is None:
            self._structs()
        stream = _raw_stream(self.device.index)
        if not fused:
            self._chain_armed = None
            rc, rst, rout = self._refs
            with torch.cuda.device(self.device):
                _cabi.check(self.lib.mg_step(rc, self.num_envs, rst, actions.data_ptr(), rout, stream), "mg_step")
        else:
            # one prepared launch per variant (mg_step_plan_*): validation, planning and knob lookups happen once
            if self._static_state is not False and self._static_ok():
                key = 1  # MG_FLAG_STATIC_GRID: a plain launch of the static-grid kernel (chained or not)
                self._chain_armed = None
            elif chained and self.one_hot is None:  # (the one-hot variants are plain launches)
                # head of a chain unless the engine's previous operation was a chained step on this stream
                key = 2 if self._chain_armed == stream else 3
                self._chain_armed = stream
            else:
                key = 0
                self._chain_armed = None
            plan = self._plans.get(key)
            ptr = actions.data_ptr()
            if plan is None or (ptr & 15):
                plan = self._make_plan(key, ptr)
            if _current_device() == self.device.index:  # (a device guard costs more than the launch)
                rc_ = self._plan_run(plan.handle, ptr, stream)
            else:
                with torch.cuda.device(self.device):
                    rc_ = self._plan_run(plan.handle, ptr, stream)
            if rc_:
                _cabi.check(rc_, "mg_step_obs")
            if self._fresh is not None:  # a new layout for every env that this step finished (see enable_fresh_layouts)
                self._chain_armed = None
                with torch.cuda.device(self.device):
                    self._refresh_done(stream)
        if self._views is None:
            self._views = (self.obs, self.reward, self.terminated, self.truncated)
        return self._views

    def _make_plan(self, key: int, actions_ptr: int):
        """mg_step_plan_create for variant `key` (0 plain, 1 static-grid, 2 chained, 3 chain head). A misaligned
        actions pointer gets a throw-away plan of its own (the planning depends on the alignment)."""
        cfg_ref = (self._refs[0], self._static_cfg[1], self._chained_cfg[1], self._chained_cfg[3])[key]
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.mg_step_plan_create(cfg_ref, self.num_envs, self._refs[1], actions_ptr, self._refs[2],
                                                     C.byref(handle)), "mg_step_plan_create")
        plan = _Plan(self.lib, handle)
        if not actions_ptr & 15:
            self._plans[key] = plan
        return plan

    def rollout(self, actions: torch.Tensor, out: dict | None = None) -> dict:
        """mg_rollout: T = actions.shape[0] consecutive fused steps in ONE launch on an open-loop
        action tape (T, E, n) int8. Bit-identical to T calls of `step(actions[t])`; returns (and
        optionally reuses, `out=`) a dict of device tensors with a leading T axis: obs
        (T,E,n,V,V,3) view of obs_buf (T,E,n,stride), direction (T,E,n), reward, terminated,
        truncated (T,E). Asynchronous on the current CUDA stream."""
        self._chain_armed = None
        if (actions.dtype != torch.int8 or not actions.is_contiguous() or actions.device != self.device
                or actions.dim() != 3 or tuple(actions.shape[1:]) != (self.num_envs, self.cfg.num_agents)):
            raise TypeError("actions must be a contiguous int8 CUDA tensor of shape (T, num_envs, n)")
        T, E, n, V = int(actions.shape[0]), self.num_envs, self.cfg.num_agents, self.cfg.view_size
        if out is None:
            e = lambda shape, dt: torch.empty(shape, dtype=dt, device=self.device)  # noqa: E731
            out = dict(obs_buf=e((T, E, n, self.obs_stride), torch.int8), direction=e((T, E, n), torch.int8),
                       reward=e((T, E, n), torch.float64), terminated=e((T, E, n), torch.uint8),
                       truncated=e((T, E), torch.uint8))
        assert out["obs_buf"].shape[0] == T and out["obs_buf"].is_contiguous()
        c, st, _ = self._structs()
        ro = _cabi.MgRolloutOut(out["obs_buf"].data_ptr(), out["direction"].data_ptr(), out["reward"].data_ptr(),
                                out["terminated"].data_ptr(), out["truncated"].data_ptr(), self.status.data_ptr())
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.mg_rollout(C.byref(c), E, T, C.byref(st), actions.data_ptr(), C.byref(ro),
                                            self._stream()), "mg_rollout")
        out["obs"] = out["obs_buf"][..., :3 * V * V].unflatten(-1, (V, V, 3))
        return out

    def obs_features(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """mg_obs_features: the current observations as the reference's network input
        (OneHotObsWrapper + scripts/train.py:56-63 preprocess_batch): float32 (E, n, V, V, 23)."""
        E, n, V = self.num_envs, self.cfg.num_agents, self.cfg.view_size
        if getattr(self, "_dir_lut", None) is None:
            d = 2 * torch.pi * (torch.arange(4) / 4)  # exactly the reference's float32 expression, on the CPU
            self._dir_lut = torch.stack([torch.cos(d), torch.sin(d)], dim=-1).contiguous().to(self.device)
        if out is None:
            out = torch.empty((E, n, V, V, 23), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.mg_obs_features(V, E * n, self.obs_stride, self.obs_buf.data_ptr(),
                                                 self.agents.data_ptr(), 8, self._dir_lut.data_ptr(), out.data_ptr(),
                                                 self._stream()), "mg_obs_features")
        return out

    def wire_palette(self):
        """(bits, codes) of the palette wire format (mg_pack_obs_palette): `codes` (uint16, sorted) = every 9-bit cell
        code type | colour << 4 | state << 7 an observation of this batch can show -- unseen, empty, the out-of-bounds
        wall, every cell of the layout pool and of the current grids (a door in all three states: toggling changes
        only the state), everything an agent carries, and the agents themselves (colour x 4 directions) -- and `bits`
        = ceil(log2(len(codes))). Derived once per pool / injected state on the device. A code outside it (objects
        injected later without telling the engine) sets bit 3 of the status word: check_status() raises."""
        if self._palette is None:
            dev = self.device
            code = lambda w: ((w & 15) | (((w >> 8) & 7) << 4) | (((w >> 16) & 3) << 7))  # noqa: E731
            found = [torch.tensor([0, 1, 2 | (5 << 4)], dtype=torch.int64, device=dev)]  # unseen, empty, wall
            for cells in (self.pool_grid, self.cells):
                if cells is not None:
                    found.append(torch.unique(code(cells.to(torch.int64).flatten())))
            for ag in (self.pool_agents, self.agents):
                if ag is not None:
                    a = ag.to(torch.int64).reshape(-1, 8)
                    found.append(torch.unique((a[:, 4] & 15) | ((a[:, 5] & 7) << 4) | ((a[:, 6] & 3) << 7)))  # carried
                    colours = torch.unique(a[:, 7] & 7)
                    found.append((10 | (colours[:, None] << 4) | (torch.arange(4, device=dev)[None, :] << 7)).flatten())
            codes = torch.unique(torch.cat(found))
            doors = codes[(codes & 15) == 4] & 0x7f  # (type 4: every state of every door colour present)
            if doors.numel():
                codes = torch.unique(torch.cat([codes, (doors[:, None] | (torch.arange(3, device=dev)[None, :] << 7)).flatten()]))
            codes = codes.cpu().numpy().astype(np.uint16)
            bits = max(1, int(np.ceil(np.log2(len(codes)))))
            if len(codes) > 255:  # (index 0xff means "not in the palette")
                raise RuntimeError("more than 255 distinct cell values: use the 9-bit wire format")
            lut = np.full(512, 0xff, np.uint8)
            lut[codes] = np.arange(len(codes), dtype=np.uint8)
            self._palette = (bits, codes, torch.from_numpy(lut).to(dev))
        return self._palette[0], self._palette[1]

    def host_buffers(self, packed=False):
        """Pinned host mirrors used by `step_host` (allocated on first use). packed=True adds `obs_packed`
        (E, n, packed_obs_stride(V)) uint8, the compact wire format of the observations (see unpack_obs);
        packed="palette" adds `obs_palette` (E, n, packed_obs_stride(V, bits)) for the palette format."""
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)  # noqa: E731
        if self._host is None:
            self._host = dict(actions=pin(self.actions), reward=pin(self.reward), terminated=pin(self.terminated),
                              truncated=pin(self.truncated))
        if packed == "wire":
            if "wire" in self._host and self._palette is not None and self._host.get("_wire_for") is self._palette:
                return self._host  # (fast path: buffers match the current palette)
            nbytes = self.lib.mg_wire_bytes(self.cfg.view_size, self.wire_palette()[0], self.cfg.num_agents, self.num_envs)
            self._host["_wire_for"] = self._palette
            if "wire" not in self._host or self._host["wire"].numel() != nbytes:
                self._wire = torch.zeros((nbytes,), dtype=torch.uint8, device=self.device)
                self._host["wire"] = pin(self._wire)
            return self._host
        if packed == "palette":
            ps = packed_obs_stride(self.cfg.view_size, self.wire_palette()[0])
            if "obs_palette" not in self._host or self._host["obs_palette"].shape[-1] != ps:
                self._packed_pal = torch.zeros((self.num_envs, self.cfg.num_agents, ps), dtype=torch.uint8, device=self.device)
                self._host["obs_palette"] = pin(self._packed_pal)
        elif packed and "obs_packed" not in self._host:
            ps = packed_obs_stride(self.cfg.view_size)
            self._packed = torch.zeros((self.num_envs, self.cfg.num_agents, ps), dtype=torch.uint8, device=self.device)
            self._host["obs_packed"] = pin(self._packed)
        if not packed and "obs" not in self._host:
            self._host["obs"] = pin(self.obs_buf)
        return self._host

    def step_host(self, synchronize: bool = True, packed=False, actions: torch.Tensor | None = None):
        """mg_step_obs_host: actions come from, and results go to, pinned HOST buffers.

        Fill `host_buffers()['actions']` first. This is the end-to-end path a CPU-side caller of
        `env.step()` sees: H2D actions + kernel + D2H (obs, reward, terminated, truncated).
        packed=True (mg_step_obs_host_packed): the observations cross PCIe in the 9-bit-per-cell wire format
        (`obs_packed`, 56 instead of 148+ bytes per agent for V = 7; `unpack_obs` decodes it), everything else
        is unchanged. packed="palette" (mg_step_obs_host_palette): `wire_palette()[0]` bits per cell (`obs_palette`:
        32 bytes per agent on Empty-8x8 with 4 agents; `unpack_obs(h["obs_palette"], V, *eng.wire_palette())`).
        packed="wire" (mg_step_obs_host_wire): ONE device-to-host copy of `h["wire"]` = palette observations followed by
        a 16-byte record per env (reward value + counts, terminated mask, truncated) instead of n float64 + n + 1 bytes;
        `unpack_wire(h["wire"], E, n, V, *eng.wire_palette())` decodes all four results.
        """
        self._chain_armed = None
        h = self.host_buffers(packed)
        c, st, out = self._structs()
        if self._static_state is not False and self._static_ok():
            c = self._static_cfg[0]
        if actions is None:
            h_act = h["actions"].data_ptr()
        else:  # the caller's own pinned int8 (E, n) tensor instead of host_buffers()['actions'] (no staging copy)
            if (actions.dtype != torch.int8 or actions.shape != self._act_shape or not actions.is_contiguous()
                    or actions.is_cuda or not actions.is_pinned()):
                raise TypeError("actions must be a contiguous pinned int8 host tensor of shape (num_envs, n)")
            h_act = actions.data_ptr()
        if packed == "wire":  # ONE device-to-host copy: palette observations + compact per-env records (unpack_wire)
            with torch.cuda.device(self.device):
                _cabi.check(self.lib.mg_step_obs_host_wire(
                    C.byref(c), self.num_envs, C.byref(st), h_act, self.actions.data_ptr(),
                    C.byref(out), self._wire.data_ptr(), self._palette[0], self._palette[2].data_ptr(),
                    h["wire"].data_ptr(), self._stream()), "mg_step_obs_host_wire")
                if synchronize:
                    torch.cuda.current_stream(self.device).synchronize()
            return h
        key = "obs_palette" if packed == "palette" else ("obs_packed" if packed else "obs")
        hout = _cabi.MgStepOut(h[key].data_ptr(), h["reward"].data_ptr(),
                               h["terminated"].data_ptr(), h["truncated"].data_ptr(), None, None)
        with torch.cuda.device(self.device):
            if packed == "palette":
                _cabi.check(self.lib.mg_step_obs_host_palette(
                    C.byref(c), self.num_envs, C.byref(st), h_act, self.actions.data_ptr(),
                    C.byref(out), self._packed_pal.data_ptr(), self._palette[0], self._palette[2].data_ptr(),
                    C.byref(hout), self._stream()), "mg_step_obs_host_palette")
            elif packed:
                _cabi.check(self.lib.mg_step_obs_host_packed(
                    C.byref(c), self.num_envs, C.byref(st), h_act, self.actions.data_ptr(),
                    C.byref(out), self._packed.data_ptr(), C.byref(hout), self._stream()), "mg_step_obs_host_packed")
            else:
                _cabi.check(self.lib.mg_step_obs_host(
                    C.byref(c), self.num_envs, C.byref(st), h_act,
                    self.actions.data_ptr(), C.byref(out), C.byref(hout), self._stream()),
                    "mg_step_obs_host")
            if synchronize:
                torch.cuda.current_stream(self.device).synchronize()
        return h

    def check_status(self) -> None:
        """Raise ValueError if any kernel saw an action outside 0..6 (base.py:473-474). Syncs."""
        st = int(self.status.item())
        if st & 2:
            self.status.zero_()
            raise RecursionError("rejection sampling failed in place_obj")  # base.py:640-641
        if st & 16:
            self.status.zero_()
            raise RuntimeError("host wire: a step's rewards are not sums of one value per env; use packed='palette'")
        if st & 8:
            self.status.zero_()
            raise RuntimeError("palette wire format: an observation holds a cell value outside wire_palette() "
                               "(state injected without load_state?); use packed=True")
        if st & 5:
            self.status.zero_()
            if st & 4:
                raise RuntimeError("MG_FLAG_STATIC_GRID promise violated: an agent left the grid or carries an object")
            raise ValueError("Unknown action")

    def bytes_per_step(self, packed=False) -> dict:
        """h2d / d2h bytes of one `step_host` call."""
        E, n = self.num_envs, self.cfg.num_agents
        V = self.cfg.view_size
        if packed == "wire":
            return dict(h2d=E * n, d2h=int(self.lib.mg_wire_bytes(V, self.wire_palette()[0], n, E)))
        obs = (packed_obs_stride(V, self.wire_palette()[0]) if packed == "palette"
               else packed_obs_stride(V) if packed else self.obs_stride)
        return dict(h2d=E * n, d2h=E * n * obs + E * n * 8 + E * n + E)
