"""The static-grid path (MG_FLAG_STATIC_GRID, multigrid_b200/csrc/mg_static.cuh): memoised per-(x, y, dir)
views + agent overlay instead of the per-env cell gather. Checked against the fixtures recorded from the
unmodified reference (every Empty-family rollout) and against the C oracle on random static layouts with
walls, floor, goal and lava, every flag, ragged batches -- on the CPU build of the kernel's phase functions
(tests/hostsim) here, on the GPU through the C ABI under `-m gpu`."""
import numpy as np
import pytest

from oracle import mg_oracle as O
from oracle.c_oracle import COracle
from tests.golden_util import load_case
from tests.test_oracle_golden import cfg_from_meta

STATIC_FIXTURES = ["empty5_n1", "empty8_n2", "empty8_n4", "empty8_n4_autoreset", "empty8_n4_joint_stw",
                   "empty16_n8_v9", "empty6r_n3_nooverlap_all"]


def fixture_state(name):
    """A reference fixture as a single-layout batch: the Empty layouts of a fixture's pool are all the same
    grid (agents differ only for the random-start ids, which never auto-reset)."""
    d, meta = load_case(name)
    cfg = cfg_from_meta(meta)
    pg, pa = d["pool_grid"], O.pack_agents(d["pool_agents"])
    assert (pg == pg[:1]).all()
    if cfg.auto_reset:
        assert (pa == pa[:1]).all()
    B = meta["B"]
    st = dict(grid=d["init_grid"], agents=O.pack_agents(d["init_agents"]), pcg_state=d["pcg_state"],
              pcg_inc=d["pcg_inc"], pool_grid=pg[:1], pool_agents=pa[:1], layout_idx=np.zeros(B, np.int32))
    return d, meta, cfg, st


def check_fixture(eng, d, meta, name, T=None):
    T = meta["T"] if T is None else min(T, meta["T"])
    for t in range(T):
        obs, rew, term, trunc = eng.step(d["actions"][t])
        msg = f"{name} step {t}"
        np.testing.assert_array_equal(obs, d["obs"][t], err_msg=msg)
        assert (rew == d["reward"][t]).all(), msg  # bit-exact float64
        np.testing.assert_array_equal(term, d["terminated"][t], err_msg=msg)
        np.testing.assert_array_equal(trunc, d["truncated"][t], err_msg=msg)
        if t % 10 == 0 or t == T - 1:
            np.testing.assert_array_equal(eng.grid, d["grid"][t], err_msg=msg)
            np.testing.assert_array_equal(O.unpack_agents(eng.agents), d["agents"][t], err_msg=msg)
            np.testing.assert_array_equal(eng.step_count, d["step_count"][t], err_msg=msg)


def static_batch(cfg, B, seed):
    """One random static layout (wall ring, interior of empty / wall / floor / goal / lava) and B envs on it
    whose agents stand anywhere off the walls (some terminated), as the promise allows."""
    rng = np.random.default_rng(seed)
    W, H, n = cfg.W, cfg.H, cfg.n
    grid = np.zeros((1, W, H, 3), np.int8)
    grid[..., 0] = O.EMPTY
    r = rng.random((W, H))
    kinds = [(O.WALL, 5, 0.12), (O.FLOOR, None, 0.08), (O.GOAL, 1, 0.06), (O.LAVA, 0, 0.06)]
    lo = 0.0
    for t, c, frac in kinds:
        m = (r >= lo) & (r < lo + frac)
        grid[0, m, 0] = t
        grid[0, m, 1] = rng.integers(0, 6, m.sum()) if c is None else c
        lo += frac
    for sl in (np.s_[0, 0, :], np.s_[0, W - 1, :], np.s_[0, :, 0], np.s_[0, :, H - 1]):
        grid[sl] = (O.WALL, 5, 0)
    free = np.argwhere(grid[0, :, :, 0] != O.WALL)
    if len(free) == 0:
        grid[0, 1, 1] = (O.EMPTY, 0, 0)
        free = np.array([[1, 1]])

    def agents_for(count):
        a = np.zeros((count, n, 8), np.int8)
        pos = free[rng.integers(0, len(free), (count, n))]
        a[..., O.A_X], a[..., O.A_Y] = pos[..., 0], pos[..., 1]
        a[..., O.A_DIR] = rng.integers(0, 4, (count, n))
        a[..., O.A_CT] = O.EMPTY
        a[..., O.A_COLOR] = np.arange(n) % 6
        return a

    agents = agents_for(B)
    agents[..., O.A_TERM] = rng.random((B, n)) < 0.1
    return dict(
        grid=np.repeat(grid, B, 0), agents=agents,
        pcg_state=rng.integers(0, 2**63, (B, 2)).astype(np.uint64) * np.uint64(2) + np.uint64(1),
        pcg_inc=rng.integers(0, 2**63, (B, 2)).astype(np.uint64) * np.uint64(2) + np.uint64(1),
        pool_grid=grid, pool_agents=agents_for(1), layout_idx=np.zeros(B, np.int32),
        step_count=rng.integers(0, 5, B).astype(np.int32))


STATIC_RANDOM = [
    (0, 203, dict(W=8, H=8, n=4, V=7)),
    (1, 77, dict(W=8, H=8, n=4, V=7, auto_reset=True, max_steps=9, joint_reward=True)),
    (2, 130, dict(W=16, H=16, n=8, V=9, success_any=False, failure_any=True)),
    (3, 64, dict(W=11, H=6, n=2, V=7, allow_agent_overlap=False, auto_reset=True, max_steps=14)),
    (4, 33, dict(W=5, H=5, n=1, V=3, success_any=False)),
    (5, 95, dict(W=7, H=9, n=3, V=5, see_through_walls=True, auto_reset=True, max_steps=12)),
    (6, 50, dict(W=10, H=6, n=12, V=11, allow_agent_overlap=False, joint_reward=True, failure_any=True)),
    (7, 17, dict(W=6, H=6, n=5, V=7, auto_reset=True, max_steps=6)),
    (8, 1, dict(W=3, H=3, n=1, V=3)),
    (9, 40, dict(W=12, H=12, n=32, V=15, auto_reset=True, max_steps=9)),
    (10, 129, dict(W=9, H=9, n=2, V=7, joint_reward=True, success_any=False, allow_agent_overlap=False)),
]


def run_vs_oracle(make_engine, seed, B, kw, T=40):
    kw = dict(kw)
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    st = static_batch(cfg, B, seed)
    ora, eng = COracle(cfg, **st), make_engine(cfg, st)
    rng = np.random.default_rng(seed + 100)
    for t in range(T):
        actions = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        o2, r2, t2, tr2 = eng.step(actions)
        msg = f"seed {seed} step {t}"
        np.testing.assert_array_equal(o2, o1, err_msg=msg)
        assert (r1 == r2).all(), msg
        np.testing.assert_array_equal(t2, t1, err_msg=msg)
        np.testing.assert_array_equal(tr2, tr1, err_msg=msg)
        np.testing.assert_array_equal(eng.grid, ora.grid, err_msg=msg)
        np.testing.assert_array_equal(eng.agents, ora.agents, err_msg=msg)
        np.testing.assert_array_equal(eng.step_count, ora.step_count, err_msg=msg)
        np.testing.assert_array_equal(eng.pcg_state, ora.pcg_state, err_msg=msg)
    return eng


# ---- CPU: the kernel's phase functions lane by lane (tests/hostsim) ------------------------------------
@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("name", STATIC_FIXTURES)
def test_hostsim_static_matches_reference(name, generic):
    from tests.hostsim.sim import SimEngine
    d, meta, cfg, st = fixture_state(name)
    sim = SimEngine(cfg, static=True, generic=generic, **st)
    np.testing.assert_array_equal(sim.gen_obs(), d["obs0"])
    check_fixture(sim, d, meta, name, T=150)


@pytest.mark.parametrize("seed,B,kw", STATIC_RANDOM)
def test_hostsim_static_random_vs_c_oracle(seed, B, kw):
    from tests.hostsim.sim import SimEngine
    run_vs_oracle(lambda cfg, st: SimEngine(cfg, static=True, forced_group=(32 if seed % 3 == 0 else 0), **st),
                  seed, B, kw)


def test_static_layout_ok_rejects_what_can_change():
    from multigrid_b200.engine import static_layout_ok
    cfg = O.OracleConfig(W=8, H=8, n=2, V=7)
    st = static_batch(cfg, 1, 0)
    assert static_layout_ok(st["pool_grid"], st["pool_agents"])
    for t in (O.DOOR, O.KEY, O.BALL, O.BOX, 0):
        g = st["pool_grid"].copy()
        g[0, 3, 3] = (t, 1, 0)
        assert not static_layout_ok(g, st["pool_agents"])
    a = st["pool_agents"].copy()
    a[0, 0, O.A_CT] = O.KEY
    assert not static_layout_ok(st["pool_grid"], a)       # carries something
    a = st["pool_agents"].copy()
    a[0, 1, O.A_X] = 0
    assert not static_layout_ok(st["pool_grid"], a)       # stands in the wall ring
    a = st["pool_agents"].copy()
    a[0, 1, O.A_X] = -1
    assert not static_layout_ok(st["pool_grid"], a)       # not placed
    assert not static_layout_ok(np.repeat(st["pool_grid"], 2, 0), np.repeat(st["pool_agents"], 2, 0))  # two layouts


# ---- GPU: through the C ABI ------------------------------------------------------------------------------
def gpu_engine(cfg, st, **kw):
    from tests.gpu_adapter import GpuEngine
    return GpuEngine(cfg, **st, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["device", "host"])
@pytest.mark.parametrize("name", STATIC_FIXTURES)
def test_gpu_static_matches_reference(name, variant):
    d, meta, cfg, st = fixture_state(name)
    g = gpu_engine(cfg, st, host_path=(variant == "host"))
    np.testing.assert_array_equal(g.gen_obs(), d["obs0"])
    check_fixture(g, d, meta, name)
    assert g.eng._static_state is True, "the engine did not take the static path"


@pytest.mark.gpu
@pytest.mark.parametrize("seed,B,kw", STATIC_RANDOM)
def test_gpu_static_random_vs_c_oracle(seed, B, kw, monkeypatch):
    if seed % 3 == 0:
        monkeypatch.setenv("MG_GROUP", "32")
    if seed % 4 == 1:
        monkeypatch.setenv("MG_NO_BULK", "1")
    g = run_vs_oracle(gpu_engine, seed, B, kw)
    assert g.eng._static_state is True


@pytest.mark.gpu
def test_gpu_static_equals_general_kernel_full_size():
    """BASELINE configs[1] at full size: the static path and the general kernel (use_static = False) produce
    identical outputs and state on the same seeded rollout with auto-reset."""
    import torch
    from multigrid_b200.engine import EngineConfig, StepEngine
    import bench
    E, n = 65536, 4
    cfg = EngineConfig(width=8, height=8, num_agents=n, view_size=7, max_steps=24, auto_reset=True)
    pg, pa = bench.empty_layout(8, n)
    engines = []
    for use_static in (True, False):
        eng = StepEngine(cfg, E, "cuda:0", pg, pa)
        eng.use_static = use_static
        st, inc = bench.pcg_words(0, E)
        eng.load_state(pcg_state=st, pcg_inc=inc)
        eng.reset_from_pool()
        engines.append(eng)
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    for t in range(60):
        a = torch.randint(0, 7, (E, n), generator=gen, device="cuda:0", dtype=torch.int32).to(torch.int8)
        outs = [tuple(x.clone() for x in eng.step(a)) for eng in engines]
        for x, y in zip(*outs):
            assert torch.equal(x, y), f"step {t}"
    a, b = engines
    assert a._static_state is True and b._static_ok() is False
    for name in ("cells", "agents", "step_count", "pcg_state", "layout_idx"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    a.check_status()
    b.check_status()


@pytest.mark.gpu
def test_gpu_static_falls_back_when_the_promise_breaks():
    """Injected state that is not the layout (or agents that carry / stand in walls) -> general kernel."""
    cfg = O.OracleConfig(W=8, H=8, n=4, V=7)
    st = static_batch(cfg, 40, 3)
    st["grid"] = st["grid"].copy()
    st["grid"][7, 3, 3] = (O.KEY, 2, 0)
    g = gpu_engine(cfg, st)
    ora = COracle(cfg, **st)
    rng = np.random.default_rng(0)
    for t in range(20):
        actions = rng.integers(0, 7, size=(40, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        o2, r2, t2, tr2 = g.step(actions)
        np.testing.assert_array_equal(o2, o1, err_msg=f"step {t}")
    assert g.eng._static_state is False
    np.testing.assert_array_equal(g.grid, ora.grid)
    np.testing.assert_array_equal(g.agents, ora.agents)
