"""The packed observation wire format of the host-buffer path (mg_pack_obs / mg_step_obs_host_packed,
include/multigrid_b200.h): 9 bits per cell, decoded by multigrid_b200.engine.unpack_obs. Lossless, so the decoded
images must equal the reference's / the oracle's observations bit for bit."""
import numpy as np
import pytest

from oracle import mg_oracle as O
from oracle.c_oracle import COracle
from tests.randstate import random_batch


def test_unpack_obs_against_a_bigint_encoder():
    from multigrid_b200.engine import packed_obs_stride, unpack_obs
    rng = np.random.default_rng(0)
    for V in (3, 5, 7, 9, 11, 15):
        img = np.stack([rng.integers(0, 11, (3, 2, V, V)), rng.integers(0, 6, (3, 2, V, V)),
                        rng.integers(0, 4, (3, 2, V, V))], -1).astype(np.int8)
        ps = packed_obs_stride(V)
        assert ps % 8 == 0 and ps * 8 >= 9 * V * V
        packed = np.zeros((3, 2, ps), np.uint8)
        for i in range(3):
            for j in range(2):
                v = 0
                for c, (t, col, st) in enumerate(img[i, j].reshape(-1, 3)):
                    v |= (int(t) | int(col) << 4 | int(st) << 7) << (9 * c)
                packed[i, j] = np.frombuffer(v.to_bytes(ps, "little"), np.uint8)
        np.testing.assert_array_equal(unpack_obs(packed, V), img)


def test_packed_stride_matches_the_library():
    from multigrid_b200 import _cabi
    from multigrid_b200.engine import packed_obs_stride
    lib = _cabi.load()
    for V in (3, 5, 7, 9, 11, 13, 15):
        assert lib.mg_packed_obs_stride(V) == packed_obs_stride(V)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,B,kw", [
    (0, 300, dict(W=8, H=8, n=4, V=7)),
    (1, 77, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True)),
    (2, 40, dict(W=16, H=16, n=8, V=9)),
    (3, 33, dict(W=9, H=7, n=3, V=5, see_through_walls=True)),
    (4, 17, dict(W=12, H=12, n=5, V=11, auto_reset=True, max_steps=9)),
])
def test_gpu_packed_host_path_vs_c_oracle(seed, B, kw):
    """mg_step_obs_host_packed on random soups: decoded observations, rewards, terminations == C oracle."""
    import torch
    from multigrid_b200.engine import unpack_obs
    from tests.gpu_adapter import GpuEngine
    kw = dict(kw)
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    st = random_batch(cfg, B, seed)
    ora, g = COracle(cfg, **st), GpuEngine(cfg, **st)
    rng = np.random.default_rng(seed)
    for t in range(12):
        actions = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        h = g.eng.host_buffers(packed=True)
        h["actions"].copy_(torch.from_numpy(actions))
        h = g.eng.step_host(packed=True)
        np.testing.assert_array_equal(unpack_obs(h["obs_packed"], cfg.V), o1, err_msg=f"step {t}")
        assert (h["reward"].numpy() == r1).all()
        np.testing.assert_array_equal(h["terminated"].numpy(), t1)
        np.testing.assert_array_equal(h["truncated"].numpy(), tr1)
        # the device-side observations are untouched by the packing
        np.testing.assert_array_equal(g._obs(g.eng.obs_buf), o1)


@pytest.mark.gpu
def test_gpu_packed_host_path_static_fixture():
    """The static-grid path (16-byte observation slots) through the packed host path, against the reference."""
    import torch
    from multigrid_b200.engine import unpack_obs
    from tests.gpu_adapter import GpuEngine
    from tests.test_static_path import fixture_state
    d, meta, cfg, st = fixture_state("empty8_n4")
    g = GpuEngine(cfg, **st)
    for t in range(60):
        h = g.eng.host_buffers(packed=True)
        h["actions"].copy_(torch.from_numpy(np.ascontiguousarray(d["actions"][t], dtype=np.int8)))
        h = g.eng.step_host(packed=True)
        np.testing.assert_array_equal(unpack_obs(h["obs_packed"], cfg.V), d["obs"][t], err_msg=f"step {t}")
        assert (h["reward"].numpy() == d["reward"][t]).all()
    assert g.eng._static_state is True and g.eng.obs_stride == 160
