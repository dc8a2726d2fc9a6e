"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol
that include/multigrid_b200.h declares (no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from multigrid_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "multigrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert {"mg_gen_obs", "mg_step", "mg_step_obs", "mg_step_obs_host"} <= set(names)
    raw = C.CDLL(_cabi.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert set(names) == set(_cabi.EXPORTS), "ctypes binding and header disagree"


def test_abi_version_and_pure_helpers(lib):
    assert lib.mg_abi_version() == _cabi.ABI_VERSION
    for v in (3, 5, 7, 9, 15):
        assert lib.mg_obs_agent_stride(v) == _cabi.obs_agent_stride(v) >= 3 * v * v
        assert lib.mg_obs_agent_stride(v) % 4 == 0
    assert lib.mg_launch_count() == 0
    assert b"aligned" in lib.mg_error_string(-2)


def test_struct_sizes_match_header():
    assert C.sizeof(_cabi.MgConfig) == 44
    assert C.sizeof(_cabi.MgState) == 96  # 12 pointers (static_obs: ABI v8)
    assert C.sizeof(_cabi.MgRolloutOut) == 48
    assert C.sizeof(_cabi.MgStepOut) == 48  # 6 pointers (one_hot: ABI v10)


def test_argument_validation_needs_no_gpu(lib):
    c = _cabi.MgConfig(8, 8, 2, 4, 100, 0, 0, 148, 0, 1)  # even view size
    assert lib.mg_gen_obs(C.byref(c), 1, None, None, None, None) == -1
    c = _cabi.MgConfig(8, 8, 2, 7, 100, 0, 0, 146, 0, 1)  # stride too small
    assert lib.mg_gen_obs(C.byref(c), 1, None, None, None, None) == -1


def test_engine_fails_loudly_without_cuda():
    import torch
    from multigrid_b200.engine import EngineConfig, StepEngine
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        StepEngine(EngineConfig(8, 8, 2), 4)


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.EngineLibraryError):
        _cabi.load()
