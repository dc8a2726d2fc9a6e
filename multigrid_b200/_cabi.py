"""ctypes binding of the C ABI declared in include/multigrid_b200.h.

The shared library is built in-tree by `multigrid_b200.build` (nvcc, sm_100a). There is NO
fallback: if the library is missing or fails to load, importing the engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MG_LIB") or os.path.join(PKG_DIR, "_lib", "libmultigrid_b200.so")  # MG_LIB: diagnostics builds

# MgConfig.flags (include/multigrid_b200.h)
FLAG_SEE_THROUGH_WALLS = 0x01
FLAG_ALLOW_OVERLAP = 0x02
FLAG_JOINT_REWARD = 0x04
FLAG_SUCCESS_ANY = 0x08
FLAG_FAILURE_ANY = 0x10
FLAG_AUTO_RESET = 0x20
FLAG_STREAM_STATE = 0x40
FLAG_CHAINED = 0x80
FLAG_CHAIN_HEAD = 0x100
FLAG_STATIC_GRID = 0x200

HOOK_NONE = 0
HOOK_BLOCKED_UNLOCK_PICKUP = 1
HOOK_RED_BLUE_DOORS = 2
HOOK_LOCKED_HALLWAY = 3

ABI_VERSION = 11
MAX_VIEW = 15
MAX_AGENTS = 32


class MgConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("num_agents", C.c_int32),
        ("view_size", C.c_int32), ("max_steps", C.c_int32), ("flags", C.c_uint32),
        ("hook", C.c_int32), ("obs_agent_stride", C.c_int32), ("num_layouts", C.c_int32),
        ("layout_stride", C.c_int32), ("hook_param", C.c_int32),
    ]


class MgState(C.Structure):
    _fields_ = [
        ("grid", C.c_void_p), ("agents", C.c_void_p), ("step_count", C.c_void_p),
        ("pcg_state", C.c_void_p), ("pcg_inc", C.c_void_p), ("layout_idx", C.c_void_p),
        ("pool_grid", C.c_void_p), ("pool_agents", C.c_void_p), ("hook_state", C.c_void_p),
        ("pool_rep", C.c_void_p), ("chain", C.c_void_p), ("static_obs", C.c_void_p),
    ]


class MgStepOut(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p),
        ("truncated", C.c_void_p), ("status", C.c_void_p), ("one_hot", C.c_void_p),
    ]


class MgLayoutGen(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("params", C.c_int32 * 4), ("rng_state", C.c_void_p), ("rng_inc", C.c_void_p),
        ("rng_buf", C.c_void_p), ("order_buf", C.c_void_p), ("info", C.c_void_p),
    ]


LAYOUT_EMPTY_RANDOM, LAYOUT_BUP, LAYOUT_RED_BLUE_DOORS, LAYOUT_LOCKED_HALLWAY, LAYOUT_PLAYGROUND = 1, 2, 3, 4, 5


class MgRolloutOut(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("direction", C.c_void_p), ("reward", C.c_void_p),
        ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("status", C.c_void_p),
    ]


EXPORTS = {
    "mg_abi_version": (C.c_int, []),
    "mg_error_string": (C.c_char_p, [C.c_int]),
    "mg_obs_agent_stride": (C.c_int32, [C.c_int32]),
    "mg_launch_count": (C.c_int64, []),
    "mg_static_obs_stride": (C.c_int32, [C.c_int32]),
    "mg_static_obs_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "mg_build_static_obs": (C.c_int, [C.POINTER(MgConfig), C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_debug_set_trace": (None, [C.c_void_p]),
    "mg_cells_per_env": (C.c_int64, [C.c_int32, C.c_int32]),
    "mg_pack_grid": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_full_obs": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p]),
    "mg_one_hot": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_one_hot_cells": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_gen_layouts_empty_random": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_gen_layouts_red_blue_doors": (C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 7),
    "mg_gen_layouts_locked_hallway": (C.c_int, [C.c_int32] * 5 + [C.c_int64] + [C.c_void_p] * 7),
    "mg_gen_layouts_playground": (C.c_int, [C.c_int32] * 4 + [C.c_int64] + [C.c_void_p] * 10),
    "mg_gen_layouts_bup": (C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 11),
    "mg_obs_features": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "mg_unpack_grid": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_gen_obs": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p]),
    "mg_step": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p,
                          C.POINTER(MgStepOut), C.c_void_p]),
    "mg_step_obs": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p,
                              C.POINTER(MgStepOut), C.c_void_p]),
    "mg_rollout": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.c_int32, C.POINTER(MgState), C.c_void_p,
                             C.POINTER(MgRolloutOut), C.c_void_p]),
    "mg_reset_where": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p, C.c_void_p]),
    "mg_refresh_done_layouts": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.POINTER(MgLayoutGen),
                                          C.c_void_p, C.c_void_p]),
    "mg_step_plan_create": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p,
                                      C.POINTER(MgStepOut), C.POINTER(C.c_void_p)]),
    "mg_step_plan_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_step_plan_destroy": (None, [C.c_void_p]),
    "mg_packed_obs_stride": (C.c_int32, [C.c_int32]),
    "mg_packed_obs_stride_bits": (C.c_int32, [C.c_int32, C.c_int32]),
    "mg_wire_record_bytes": (C.c_int32, [C.c_int32]),
    "mg_wire_obs_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "mg_wire_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "mg_step_obs_host_wire": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p, C.c_void_p,
                                        C.POINTER(MgStepOut), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_pack_obs_palette": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "mg_step_obs_host_palette": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p, C.c_void_p,
                                           C.POINTER(MgStepOut), C.c_void_p, C.c_int32, C.c_void_p,
                                           C.POINTER(MgStepOut), C.c_void_p]),
    "mg_pack_obs": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_step_obs_host_packed": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p, C.c_void_p,
                                          C.POINTER(MgStepOut), C.c_void_p, C.POINTER(MgStepOut), C.c_void_p]),
    "mg_step_obs_host": (C.c_int, [C.POINTER(MgConfig), C.c_int64, C.POINTER(MgState), C.c_void_p,
                                   C.c_void_p, C.POINTER(MgStepOut), C.POINTER(MgStepOut),
                                   C.c_void_p]),
}


class EngineLibraryError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA engine library (once). Raises EngineLibraryError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineLibraryError(
            f"{LIB_PATH} not found: build it with `python -m multigrid_b200.build` "
            "(nvcc, sm_100a). multigrid_b200 has no CPU or PyTorch fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # e.g. libcudart missing
        raise EngineLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = restype, argtypes
    if lib.mg_abi_version() != ABI_VERSION:
        raise EngineLibraryError(
            f"ABI mismatch: library {lib.mg_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mg_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def obs_agent_stride(view_size: int) -> int:
    return (3 * view_size * view_size + 3) & ~3
