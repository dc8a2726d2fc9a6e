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


# ---- palette wire format (mg_pack_obs_palette / mg_step_obs_host_palette, ABI v11) -------------------------------

def test_unpack_obs_palette_against_a_bigint_encoder():
    from multigrid_b200.engine import packed_obs_stride, unpack_obs
    rng = np.random.default_rng(1)
    for V, K in ((3, 2), (7, 20), (7, 33), (9, 28), (5, 256), (15, 100)):
        codes = np.sort(rng.choice(512, size=K, replace=False)).astype(np.uint16)
        bits = max(1, int(np.ceil(np.log2(K))))
        idx = rng.integers(0, K, (3, 2, V * V))
        ps = packed_obs_stride(V, bits)
        assert ps % 8 == 0 and ps * 8 >= bits * V * V
        packed = np.zeros((3, 2, ps), np.uint8)
        for i in range(3):
            for j in range(2):
                v = 0
                for c, q in enumerate(idx[i, j]):
                    v |= int(q) << (bits * c)
                packed[i, j] = np.frombuffer(v.to_bytes(ps, "little"), np.uint8)
        code = codes[idx].astype(np.int64)
        img = np.stack([code & 15, (code >> 4) & 7, code >> 7], -1).astype(np.int8).reshape(3, 2, V, V, 3)
        np.testing.assert_array_equal(unpack_obs(packed, V, bits, codes), img)


def test_palette_stride_matches_the_library():
    from multigrid_b200 import _cabi
    from multigrid_b200.engine import packed_obs_stride
    lib = _cabi.load()
    for V in (3, 7, 9, 15):
        for bits in range(1, 9):
            assert lib.mg_packed_obs_stride_bits(V, bits) == packed_obs_stride(V, bits)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,B,kw", [
    (0, 300, dict(W=8, H=8, n=4, V=7)),
    (1, 77, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True)),
    (2, 40, dict(W=16, H=16, n=8, V=9)),
    (3, 33, dict(W=9, H=7, n=3, V=5, see_through_walls=True)),
    (4, 17, dict(W=12, H=12, n=5, V=11, auto_reset=True, max_steps=9)),
])
def test_gpu_palette_host_path_vs_c_oracle(seed, B, kw):
    """mg_step_obs_host_palette on random object soups (doors toggled, objects carried and dropped, auto-reset): the
    palette derived from the pool and the injected state covers every step; decoded observations == C oracle."""
    import torch
    from multigrid_b200.engine import unpack_obs
    from tests.gpu_adapter import GpuEngine
    kw = dict(kw)
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    st = random_batch(cfg, B, seed)
    ora, g = COracle(cfg, **st), GpuEngine(cfg, **st)
    bits, codes = g.eng.wire_palette()
    assert 1 <= bits <= 8 and len(codes) <= 1 << bits
    rng = np.random.default_rng(seed)
    for t in range(20):
        actions = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        h = g.eng.host_buffers(packed="palette")
        h["actions"].copy_(torch.from_numpy(actions))
        h = g.eng.step_host(packed="palette")
        np.testing.assert_array_equal(unpack_obs(h["obs_palette"], cfg.V, bits, codes), o1, err_msg=f"step {t}")
        assert (h["reward"].numpy() == r1).all()
        np.testing.assert_array_equal(h["terminated"].numpy(), t1)
        np.testing.assert_array_equal(h["truncated"].numpy(), tr1)
    g.eng.check_status()


@pytest.mark.gpu
def test_gpu_palette_static_fixture_is_five_bits():
    """Empty-8x8 with 4 agents: unseen, empty, wall, goal + 16 agent encodings = 20 cell values -> 5 bits, 32 bytes
    per view; decoded observations equal the reference's."""
    import torch
    from multigrid_b200.engine import unpack_obs
    from tests.gpu_adapter import GpuEngine
    from tests.test_static_path import fixture_state
    d, meta, cfg, st = fixture_state("empty8_n4")
    g = GpuEngine(cfg, **st)
    bits, codes = g.eng.wire_palette()
    assert bits == 5 and len(codes) == 20
    for t in range(60):
        h = g.eng.host_buffers(packed="palette")
        h["actions"].copy_(torch.from_numpy(np.ascontiguousarray(d["actions"][t], dtype=np.int8)))
        h = g.eng.step_host(packed="palette")
        assert h["obs_palette"].shape == (meta["B"], 4, 32)
        np.testing.assert_array_equal(unpack_obs(h["obs_palette"], cfg.V, bits, codes), d["obs"][t], err_msg=f"step {t}")
        assert (h["reward"].numpy() == d["reward"][t]).all()
    g.eng.check_status()


@pytest.mark.gpu
def test_gpu_palette_flags_a_value_it_does_not_hold():
    """A cell value that appears behind the engine's back (grid written directly) is outside the palette: the pack
    kernel flags it and check_status() raises instead of delivering a wrong image."""
    import torch
    from tests.gpu_adapter import GpuEngine
    cfg = O.OracleConfig(W=8, H=8, n=2, V=7, max_steps=50)
    st = random_batch(cfg, 20, 3)
    st["grid"][:] = st["grid"][:1]
    st["grid"][:, 1:-1, 1:-1] = (1, 0, 0)  # empty interiors
    g = GpuEngine(cfg, **st)
    bits, codes = g.eng.wire_palette()
    t, c = next((t, c) for t in (9, 3, 6, 5, 7) for c in range(6) if (t | c << 4) not in set(codes.tolist()))
    g.eng.cells[:, 1:8, 1:8] = t | (c << 8)  # an object the palette has never seen, written without load_state
    h = g.eng.host_buffers(packed="palette")
    h["actions"].fill_(6)
    g.eng.step_host(packed="palette")
    assert (g.eng.obs[..., 0] == t).any()
    with pytest.raises(RuntimeError):
        g.eng.check_status()


# ---- the host wire: palette observations + compact env records in one copy (mg_step_obs_host_wire) ---------------

def test_wire_env_records_round_trip_on_the_cpu():
    """wire_env_record (the kernel's function, compiled for the host) + engine.unpack_wire: rewards that are 0 or
    k-fold sums of one value per env come back bit for bit; anything else is reported."""
    import ctypes as C
    from multigrid_b200.engine import unpack_wire, wire_record_bytes
    from tests.hostsim import sim as S
    rng = np.random.default_rng(5)
    for n in (1, 2, 4, 8, 9, 31):
        E, V, bits = 64, 3, 2
        value = 1 - 0.9 * (rng.integers(1, 200, E) / 256)
        # per env: individual rewards (each agent 0 or 1 x value) or joint ones (every agent the same k x value, k = the
        # doors unlocked in that step under the LockedHallway hook), or a mix of 1 x and k x
        k_env = rng.integers(1, 4, (E, 1))
        counts = np.where(rng.random((E, n)) < 0.3, np.where(rng.random((E, 1)) < 0.5, 1, k_env), 0)
        counts[:, 0] = np.where(counts.max(1) > 1, 1, counts[:, 0])  # (a 1 x agent next to the k x ones)
        reward = np.zeros((E, n))
        for k in range(1, 4):
            reward = np.where(counts >= k, reward + value[:, None], reward)
        term = (rng.random((E, n)) < 0.4).astype(np.uint8)
        trunc = (rng.random(E) < 0.2).astype(np.uint8)
        rb = wire_record_bytes(n)
        rec = S.aligned((E, rb), np.uint8)
        rc = S.lib().sim_wire_env_records(C.c_int(n), C.c_int64(E), S._p(S.aligned_copy(reward, np.float64)),
                                          S._p(S.aligned_copy(term, np.uint8)), S._p(S.aligned_copy(trunc, np.uint8)), S._p(rec))
        assert rc == 0
        ps = 8  # (V = 3, 2 bits: 18 bits -> one word per agent)
        obs_bytes = (E * n * ps + 15) & ~15
        wire = np.concatenate([np.zeros(obs_bytes, np.uint8), rec.reshape(-1)])
        _, r2, t2, tr2 = unpack_wire(wire, E, n, V, bits, np.zeros(4, np.uint16))
        assert (r2 == reward).all() and r2.dtype == np.float64
        np.testing.assert_array_equal(t2, term.astype(bool))
        np.testing.assert_array_equal(tr2, trunc.astype(bool))
    # two different values in one env cannot be represented: reported
    reward = np.array([[0.5, 0.3]])
    rec = S.aligned((1, wire_record_bytes(2)), np.uint8)
    assert S.lib().sim_wire_env_records(C.c_int(2), C.c_int64(1), S._p(S.aligned_copy(reward, np.float64)),
                                        S._p(S.aligned((1, 2), np.uint8)), S._p(S.aligned((1,), np.uint8)), S._p(rec)) == 1


def test_wire_sizes_match_the_library():
    from multigrid_b200 import _cabi
    from multigrid_b200.engine import packed_obs_stride, wire_record_bytes
    lib = _cabi.load()
    for n in (1, 4, 8, 9, 16, 31):
        assert lib.mg_wire_record_bytes(n) == wire_record_bytes(n)
        for V, bits, E in ((7, 5, 1000), (9, 6, 33), (3, 1, 1)):
            ob = (E * n * packed_obs_stride(V, bits) + 15) & ~15
            assert lib.mg_wire_obs_bytes(V, bits, n, E) == ob
            assert lib.mg_wire_bytes(V, bits, n, E) == ob + E * wire_record_bytes(n)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["empty8_n4_autoreset", "bup_n2_teleport", "lh4_n3_nojoint", "lh2_n2", "rbd_n2", "soup_0", "soup_autoreset"])
def test_gpu_host_wire_matches_reference_fixtures(name):
    """mg_step_obs_host_wire against rollouts of the unmodified reference, decoded by unpack_wire: images, float64
    rewards (incl. the LockedHallway rewards that add up), terminations, truncations."""
    import torch
    from multigrid_b200.engine import unpack_wire
    from tests.golden_util import ROLLOUT_CASES, load_case
    from tests.gpu_adapter import GpuEngine
    from tests.test_oracle_golden import cfg_from_meta
    if name not in ROLLOUT_CASES:
        pytest.skip("fixture not present")
    d, meta = load_case(name)
    cfg = cfg_from_meta(meta)
    B, T, J = meta["B"], meta["T"], meta["pool_J"]
    g = GpuEngine(cfg, d["init_grid"], O.pack_agents(d["init_agents"]), d["pcg_state"], d["pcg_inc"],
                  pool_grid=d["pool_grid"], pool_agents=O.pack_agents(d["pool_agents"]), layout_idx=np.arange(B) * J)
    bits, codes = g.eng.wire_palette()
    for t in range(T):
        h = g.eng.host_buffers(packed="wire")
        h["actions"].copy_(torch.from_numpy(np.ascontiguousarray(d["actions"][t], dtype=np.int8)))
        h = g.eng.step_host(packed="wire")
        img, rew, term, trunc = unpack_wire(h["wire"], B, cfg.n, cfg.V, bits, codes)
        msg = f"{name} step {t}"
        np.testing.assert_array_equal(img, d["obs"][t], err_msg=msg)
        assert (rew == d["reward"][t]).all(), msg
        np.testing.assert_array_equal(term, d["terminated"][t].astype(bool), err_msg=msg)
        np.testing.assert_array_equal(trunc, d["truncated"][t].astype(bool), err_msg=msg)
    g.eng.check_status()


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", ["2", "4", "8"])
@pytest.mark.parametrize("seed,B,kw", [
    (0, 1000, dict(W=8, H=8, n=4, V=7)),
    (1, 777, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True, auto_reset=True, max_steps=9)),
    (2, 600, dict(W=9, H=7, n=3, V=5, see_through_walls=True)),
])
def test_gpu_host_wire_chunked_pipeline(seed, B, kw, chunks, monkeypatch):
    """MG_WIRE_CHUNKS: the envs are stepped and packed in slices (multiples of 256 envs, ragged last slice) on the
    caller's stream while a side stream copies finished slices; results and state equal the C oracle's."""
    import torch
    from multigrid_b200.engine import unpack_wire
    from tests.gpu_adapter import GpuEngine
    monkeypatch.setenv("MG_WIRE_CHUNKS", chunks)
    kw = dict(kw)
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    st = random_batch(cfg, B, seed)
    ora, g = COracle(cfg, **st), GpuEngine(cfg, **st)
    bits, codes = g.eng.wire_palette()
    rng = np.random.default_rng(seed)
    for t in range(12):
        actions = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        h = g.eng.host_buffers(packed="wire")
        h["actions"].copy_(torch.from_numpy(actions))
        h = g.eng.step_host(packed="wire")
        img, rew, term, trunc = unpack_wire(h["wire"], B, cfg.n, cfg.V, bits, codes)
        np.testing.assert_array_equal(img, o1, err_msg=f"step {t}")
        assert (rew == r1).all()
        np.testing.assert_array_equal(term, t1.astype(bool))
        np.testing.assert_array_equal(trunc, tr1.astype(bool))
    g.eng.check_status()
    P_assert_same(g, ora)


def P_assert_same(a, b):
    np.testing.assert_array_equal(a.grid, b.grid)
    np.testing.assert_array_equal(a.agents, b.agents)
    np.testing.assert_array_equal(a.step_count, b.step_count)
    np.testing.assert_array_equal(a.pcg_state, b.pcg_state)
