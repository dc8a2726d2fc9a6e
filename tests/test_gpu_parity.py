"""GPU parity tests proper: the sm_100a kernels, called through the C ABI, against
(1) fixtures recorded from the unmodified reference, (2) the C oracle on seeded random states,
(3) the BASELINE.json configurations at full size (direct diff against the C oracle plus
size-independent properties). Bit-exact everywhere, including the float64 rewards."""
import os

import numpy as np
import pytest

from oracle import mg_oracle as O
from oracle.c_oracle import COracle
from tests.golden_util import ROLLOUT_CASES, load_case, GOLDEN_DIR
from tests.randstate import random_batch
from tests.test_oracle_golden import cfg_from_meta

pytestmark = pytest.mark.gpu

NTHREADS = max(1, len(os.sched_getaffinity(0)))


def GpuEngine(*a, **k):
    from tests.gpu_adapter import GpuEngine as G
    return G(*a, **k)


def assert_same(a, b, msg):
    np.testing.assert_array_equal(a.grid, b.grid, err_msg=msg)
    np.testing.assert_array_equal(a.agents, b.agents, err_msg=msg)
    np.testing.assert_array_equal(a.step_count, b.step_count, err_msg=msg)
    np.testing.assert_array_equal(a.pcg_state, b.pcg_state, err_msg=msg)
    np.testing.assert_array_equal(a.layout_idx, b.layout_idx, err_msg=msg)


@pytest.mark.parametrize("variant", ["device", "host", "split"])
@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_rollout_matches_reference(name, variant):
    d, meta = load_case(name)
    cfg = cfg_from_meta(meta)
    if variant == "split" and (cfg.hook or cfg.auto_reset):
        pytest.skip("split step/gen_obs is only equivalent without post-hook / auto-reset")
    B, T, J = meta["B"], meta["T"], meta["pool_J"]
    kw = dict(device={}, host=dict(host_path=True), split=dict(fused=False))[variant]
    g = GpuEngine(cfg, d["init_grid"], O.pack_agents(d["init_agents"]), d["pcg_state"],
                  d["pcg_inc"], pool_grid=d["pool_grid"],
                  pool_agents=O.pack_agents(d["pool_agents"]), layout_idx=np.arange(B) * J, **kw)
    np.testing.assert_array_equal(g.gen_obs(), d["obs0"])
    for t in range(T):
        obs, rew, term, trunc = g.step(d["actions"][t])
        msg = f"{name} step {t}"
        np.testing.assert_array_equal(obs, d["obs"][t], err_msg=msg)
        assert (rew == d["reward"][t]).all(), msg  # bit-exact float64
        np.testing.assert_array_equal(term, d["terminated"][t], err_msg=msg)
        np.testing.assert_array_equal(trunc, d["truncated"][t], err_msg=msg)
        if t % 10 == 0 or t == T - 1:
            np.testing.assert_array_equal(g.grid, d["grid"][t], err_msg=msg)
            np.testing.assert_array_equal(O.unpack_agents(g.agents), d["agents"][t], err_msg=msg)
            np.testing.assert_array_equal(g.agents[..., O.A_DIR], d["direction"][t], err_msg=msg)
            np.testing.assert_array_equal(g.step_count, d["step_count"][t], err_msg=msg)


def test_obs_random_injected_states():
    d = np.load(f"{GOLDEN_DIR}/obs_random.npz")
    for c in range(len(d["W"])):
        W, H, n, V = (int(d[k][c]) for k in ("W", "H", "n", "V"))
        cfg = O.OracleConfig(W=W, H=H, n=n, V=V, see_through_walls=bool(d["stw"][c]))
        grid = np.ascontiguousarray(d["grid"][c, :W, :H])[None]
        agents = O.pack_agents(d["agents"][c, :n])[None]
        z = np.zeros((1, 2), np.uint64)
        got = GpuEngine(cfg, grid, agents, z, z).gen_obs()[0]
        np.testing.assert_array_equal(got, d["obs"][c, :n, :V, :V], err_msg=f"case {c}")


SOUP = [
    (0, 1000, dict(W=8, H=8, n=4, V=7)),
    (1, 777, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True)),
    (2, 203, dict(W=9, H=13, n=5, V=9, allow_agent_overlap=False, failure_any=True)),
    (3, 515, dict(W=5, H=5, n=1, V=3, success_any=False)),
    (4, 301, dict(W=16, H=16, n=8, V=9, joint_reward=True, success_any=False)),
    (5, 203, dict(W=7, H=7, n=3, V=5, see_through_walls=True, auto_reset=True, max_steps=12)),
    (6, 99, dict(W=10, H=6, n=12, V=11, max_steps=30, auto_reset=True, layout_stride=3)),
    (7, 64, dict(W=6, H=9, n=2, V=13, allow_agent_overlap=False)),
    (8, 1, dict(W=19, H=19, n=3, V=7)),
    (9, 17, dict(W=25, H=25, n=2, V=15, auto_reset=True, max_steps=9)),
    (10, 40, dict(W=12, H=12, n=32, V=15, auto_reset=True, max_steps=9)),            # MG_MAX_AGENTS, MG_MAX_VIEW
    (11, 33, dict(W=3, H=3, n=1, V=3)),                                               # the smallest walled grid
    (12, 130, dict(W=9, H=4, n=31, V=5, allow_agent_overlap=False, joint_reward=True, hook=1)),
]


@pytest.mark.parametrize("seed,B,kw", SOUP)
def test_random_soup_vs_c_oracle(seed, B, kw):
    kw = dict(kw)
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    st = random_batch(cfg, B, seed)
    ora, g = COracle(cfg, **st), GpuEngine(cfg, **st)
    np.testing.assert_array_equal(g.gen_obs(), ora.gen_obs())
    rng = np.random.default_rng(seed + 100)
    for t in range(50):
        actions = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        o2, r2, t2, tr2 = g.step(actions)
        msg = f"step {t}"
        np.testing.assert_array_equal(o2, o1, err_msg=msg)
        assert (r1 == r2).all(), msg
        np.testing.assert_array_equal(t2, t1, err_msg=msg)
        np.testing.assert_array_equal(tr2, tr1, err_msg=msg)
        assert_same(g, ora, msg)


def empty_layout(size, n):
    """EmptyEnv._gen_grid with the default fixed start (envs/empty.py:151-170)."""
    grid = np.zeros((1, size, size, 3), np.int8)
    grid[..., 0] = O.EMPTY
    for sl in (np.s_[0, 0, :], np.s_[0, size - 1, :], np.s_[0, :, 0], np.s_[0, :, size - 1]):
        grid[sl] = (O.WALL, 5, 0)
    grid[0, size - 2, size - 2] = (O.GOAL, 1, 0)
    agents = np.zeros((1, n, 8), np.int8)
    agents[..., O.A_X] = 1
    agents[..., O.A_Y] = 1
    agents[..., O.A_CT] = O.EMPTY
    agents[..., O.A_COLOR] = np.arange(n) % 6
    return grid, agents


def seeded_pcg(B, base_seed):
    st = np.zeros((B, 2), np.uint64)
    inc = np.zeros((B, 2), np.uint64)
    m = (1 << 64) - 1
    for e in range(B):
        s = np.random.PCG64(np.random.SeedSequence(base_seed + e)).state["state"]
        st[e] = (s["state"] & m, s["state"] >> 64)
        inc[e] = (s["inc"] & m, s["inc"] >> 64)
    return st, inc


FULL = [
    ("Empty-8x8 n=4 E=65536 (BASELINE configs[1])", 8, 4, 7, 65536, 0),
    ("BUP-shaped 11x6 n=2 E=32768 (configs[2], random layouts)", None, 2, 7, 32768, 1),
    ("Empty-16x16 n=8 V=9 E=16384 (configs[3])", 16, 8, 9, 16384, 0),
]


@pytest.mark.parametrize("label,size,n,V,E,hook", FULL)
def test_full_size_configs_vs_c_oracle(label, size, n, V, E, hook):
    """BASELINE.json sizes: auto-reset rollout diffed against the C oracle every step, plus
    size-independent properties (obs self cell == carried object; checksum of obs)."""
    rng = np.random.default_rng(5)
    if size is not None:
        cfg = O.OracleConfig(W=size, H=size, n=n, V=V, max_steps=4 * size * size, auto_reset=True)
        pool_grid, pool_agents = empty_layout(size, n)
        st = dict(grid=np.repeat(pool_grid, E, 0), agents=np.repeat(pool_agents, E, 0),
                  pool_grid=pool_grid, pool_agents=pool_agents,
                  layout_idx=np.zeros(E, np.int32), step_count=np.zeros(E, np.int32))
    else:
        cfg = O.OracleConfig(W=11, H=6, n=n, V=V, max_steps=576, auto_reset=True, hook=hook,
                             joint_reward=True)
        st = random_batch(cfg, E, 9, K=4096)
    # a slice of envs starts near the step limit so truncation + auto-reset are exercised
    st["step_count"] = np.where(np.arange(E) % 7 == 0, cfg.max_steps - 3, 0).astype(np.int32)
    st["pcg_state"], st["pcg_inc"] = seeded_pcg(E, 1234) if E <= 16384 else (
        rng.integers(0, 2**63, (E, 2)).astype(np.uint64),
        rng.integers(0, 2**63, (E, 2)).astype(np.uint64) | np.uint64(1))
    ora, g = COracle(cfg, nthreads=NTHREADS, **st), GpuEngine(cfg, **st)
    np.testing.assert_array_equal(g.gen_obs(), ora.gen_obs())
    for t in range(12):
        actions = rng.integers(0, 7, size=(E, n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        o2, r2, t2, tr2 = g.step(actions)
        msg = f"{label} step {t}"
        assert np.array_equal(o2, o1), msg
        assert (r1 == r2).all(), msg
        assert np.array_equal(t2, t1) and np.array_equal(tr2, tr1), msg
        # property: every agent sees its carried object on its own cell (utils/obs.py:207)
        assert np.array_equal(o2[:, :, V // 2, V - 1, :], g.agents[:, :, O.A_CT:O.A_CS + 1]), msg
        assert int(o2.astype(np.int64).sum()) == int(o1.astype(np.int64).sum()), msg
    assert_same(g, ora, label)


KNOBS = [dict(MG_NO_BULK="1"), dict(MG_GROUP="8"), dict(MG_GROUP="32"), dict(MG_GROUP="32", MG_WPB="1"),
         dict(MG_WPB="2"), dict(MG_WPB="1", MG_NO_BULK="1"), dict(MG_GENERIC_VIEW="1"),
         dict(MG_PDL="1"), dict(MG_PDL="0"), dict(MG_L2HINT="3"), dict(MG_PDL="1", MG_L2HINT="1")]


@pytest.mark.parametrize("knobs", KNOBS, ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()))
@pytest.mark.parametrize("seed,B,kw", [SOUP[0], SOUP[1], SOUP[2], SOUP[4], SOUP[5], SOUP[6]])
def test_launch_knobs_do_not_change_results(seed, B, kw, knobs, monkeypatch):
    """Group size, warps per block, TMA-vs-plain copies and the rolled-loop view are implementation choices only."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    test_random_soup_vs_c_oracle(seed, B, kw)


@pytest.mark.parametrize("rep", range(3))
@pytest.mark.parametrize("pdl,chained", [("0", False), ("1", False), ("1", True), ("0", True)])
def test_back_to_back_launches_without_host_sync(pdl, chained, rep, monkeypatch):
    """T dependent launches enqueued on one stream with no synchronisation in between (eager and as a
    CUDA graph): with programmatic dependent launch every launch may be scheduled while its
    predecessor drains, and must still see all of the predecessor's writes. chained=True
    (MG_FLAG_CHAINED): launches are ordered env by env through the chain tickets instead of waiting
    for the whole previous grid -- every warp of launch k+1 races the other warps of launch k."""
    import torch
    monkeypatch.setenv("MG_PDL", pdl)
    cfg = O.OracleConfig(W=8, H=8, n=4, V=7, max_steps=40, auto_reset=True)
    B, T = (20000, 3000, 66000)[rep], 48  # (3 000 envs: launches far smaller than the chip, deep overlap)
    st = random_batch(cfg, B, 5 + rep)
    ora, g = COracle(cfg, nthreads=NTHREADS, **st), GpuEngine(cfg, **st)
    rng = np.random.default_rng(3 + rep)
    actions = rng.integers(0, 7, size=(2 * T, B, cfg.n)).astype(np.int8)
    tape = torch.from_numpy(actions).cuda()
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for t in range(T):
            g.eng.step(tape[t], chained=chained)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for t in range(T, 2 * T):
                g.eng.step(tape[t], chained=chained)
        graph.replay()
    torch.cuda.synchronize()
    for t in range(2 * T):
        o1, r1, t1, tr1 = ora.step(actions[t])
    np.testing.assert_array_equal(g._obs(g.eng.obs_buf), o1)
    assert (g.eng.reward.cpu().numpy() == r1).all()
    np.testing.assert_array_equal(g.eng.terminated.cpu().numpy(), t1)
    assert_same(g, ora, f"pdl={pdl} chained={chained}")
    want = 2 * T if chained else 0  # plain launches never touch the tickets
    assert (g.eng.chain[:, :2].cpu().numpy() == want).all()


@pytest.mark.parametrize("kw,B", [(dict(W=8, H=8, n=4, V=7, max_steps=30, auto_reset=True), 65536),
                                  (dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True, auto_reset=True, max_steps=14), 9000),
                                  (dict(W=16, H=16, n=8, V=9, auto_reset=True, max_steps=11), 3000)])
def test_chained_launches_rotating_engines(kw, B):
    """The bench's launch pattern: several engines (replicas) rotated on one stream, every launch chained, as
    one CUDA graph replayed twice, with unchained operations (gen_obs, a plain step, mg_rollout) in between.
    Every engine must end in the oracle's state."""
    import torch
    kw = dict(kw)
    cfg = O.OracleConfig(**kw)
    R, T = 3, 30
    sts = [random_batch(cfg, B, 50 + r) for r in range(R)]
    oras = [COracle(cfg, nthreads=NTHREADS, **st) for st in sts]
    gs = [GpuEngine(cfg, **st) for st in sts]
    rng = np.random.default_rng(9)
    actions = rng.integers(0, 7, size=(T, B, cfg.n)).astype(np.int8)
    tape = torch.from_numpy(actions).cuda()
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    done = [[] for _ in range(R)]  # action indices applied to each engine, in order
    with torch.cuda.stream(stream):
        for k in range(8):
            gs[k % R].eng.step(tape[k % T], chained=True)
            done[k % R].append(k % T)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(8, 8 + 2 * T):
                gs[k % R].eng.step(tape[k % T], chained=True)
        for rep in range(2):
            graph.replay()
            for k in range(8, 8 + 2 * T):
                done[k % R].append(k % T)
            gs[0].eng.gen_obs()                                   # unchained reader of engine 0's state
            gs[1].eng.step(tape[rep]); done[1].append(rep)        # a plain (unchained) step
            gs[2].eng.rollout(tape[:3]); done[2].extend(range(3))  # three steps in one launch
    torch.cuda.synchronize()
    for r in range(R):
        for a in done[r]:
            oras[r].step(actions[a])
        assert_same(gs[r], oras[r], f"engine {r}")
        n_launch_steps = len(done[r])
        assert (gs[r].eng.chain[:, 0].cpu().numpy() == gs[r].eng.chain[:, 1].cpu().numpy()).all()
    np.testing.assert_array_equal(gs[0].gen_obs(), oras[0].gen_obs())


@pytest.mark.parametrize("seed,B,T,kw", [
    (0, 4096, 40, dict(W=8, H=8, n=4, V=7, auto_reset=True, max_steps=9)),
    (1, 2048, 50, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True, auto_reset=True, max_steps=14)),
    (2, 300, 30, dict(W=9, H=13, n=5, V=9, allow_agent_overlap=False, failure_any=True)),
    (6, 203, 25, dict(W=10, H=6, n=12, V=11, max_steps=30, auto_reset=True, layout_stride=3)),
    (7, 77, 25, dict(W=6, H=9, n=2, V=13, allow_agent_overlap=False)),
    (8, 1024, 30, dict(W=16, H=16, n=8, V=9, joint_reward=True, auto_reset=True, max_steps=11)),
])
@pytest.mark.parametrize("knobs", [{}, dict(MG_NO_BULK="1"), dict(MG_GROUP="32")], ids=["default", "nobulk", "g32"])
def test_rollout_equals_single_steps(seed, B, T, kw, knobs, monkeypatch):
    """mg_rollout (T steps in one launch, dense random soups where cells change and envs reset):
    slice t of every output == step t of the C oracle; final state == the oracle's after T steps."""
    import torch
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    kw = dict(kw)
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    st = random_batch(cfg, B, seed)
    ora, g = COracle(cfg, nthreads=NTHREADS, **st), GpuEngine(cfg, **st)
    rng = np.random.default_rng(seed + 7)
    actions = rng.integers(-1, 7, size=(T, B, cfg.n)).astype(np.int8)
    out = g.eng.rollout(torch.from_numpy(actions).cuda())
    g.eng.check_status()
    obs = out["obs"].cpu().numpy()
    assert (out["obs_buf"][..., 3 * cfg.V * cfg.V:] == 0).all()
    dirs, rew = out["direction"].cpu().numpy(), out["reward"].cpu().numpy()
    term, trunc = out["terminated"].cpu().numpy(), out["truncated"].cpu().numpy()
    for t in range(T):
        o1, r1, t1, tr1 = ora.step(actions[t])
        msg = f"step {t}"
        np.testing.assert_array_equal(obs[t], o1, err_msg=msg)
        np.testing.assert_array_equal(dirs[t], ora.agents[..., O.A_DIR], err_msg=msg)
        assert (rew[t] == r1).all(), msg
        np.testing.assert_array_equal(term[t], t1, err_msg=msg)
        np.testing.assert_array_equal(trunc[t], tr1, err_msg=msg)
    assert_same(g, ora, "rollout")


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_rollout_matches_reference_fixtures(name):
    """mg_rollout against the rollouts recorded from the unmodified reference."""
    import torch
    d, meta = load_case(name)
    cfg = cfg_from_meta(meta)
    B, T, J = meta["B"], meta["T"], meta["pool_J"]
    g = GpuEngine(cfg, d["init_grid"], O.pack_agents(d["init_agents"]), d["pcg_state"],
                  d["pcg_inc"], pool_grid=d["pool_grid"],
                  pool_agents=O.pack_agents(d["pool_agents"]), layout_idx=np.arange(B) * J)
    out = g.eng.rollout(torch.from_numpy(np.ascontiguousarray(d["actions"][:T], dtype=np.int8)).cuda())
    g.eng.check_status()
    np.testing.assert_array_equal(out["obs"].cpu().numpy(), d["obs"][:T])
    assert (out["reward"].cpu().numpy() == d["reward"][:T]).all()
    np.testing.assert_array_equal(out["terminated"].cpu().numpy(), d["terminated"][:T])
    np.testing.assert_array_equal(out["truncated"].cpu().numpy(), d["truncated"][:T])
    np.testing.assert_array_equal(out["direction"].cpu().numpy(), d["direction"][:T])
    np.testing.assert_array_equal(g.grid, d["grid"][T - 1])
    np.testing.assert_array_equal(O.unpack_agents(g.agents), d["agents"][T - 1])


@pytest.mark.parametrize("B", [1, 15, 16, 17, 31, 32, 33, 129])
def test_ragged_batch_sizes(B):
    cfg = O.OracleConfig(W=8, H=8, n=4, V=7, max_steps=20, auto_reset=True)
    st = random_batch(cfg, B, 42)
    ora, g = COracle(cfg, **st), GpuEngine(cfg, **st)
    rng = np.random.default_rng(1)
    for t in range(30):
        actions = rng.integers(0, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        o2, r2, t2, tr2 = g.step(actions)
        assert np.array_equal(o2, o1) and (r1 == r2).all()
        assert np.array_equal(t2, t1) and np.array_equal(tr2, tr1)
    assert_same(g, ora, f"B={B}")


def test_unknown_action_raises_value_error():
    cfg = O.OracleConfig(W=8, H=8, n=2, V=7)
    st = random_batch(cfg, 4, 0)
    st["agents"][..., O.A_TERM] = 0
    g = GpuEngine(cfg, **st)
    actions = np.zeros((4, 2), np.int8)
    actions[2, 1] = 9
    with pytest.raises(ValueError):  # base.py:473-474
        g.step(actions)


def test_misaligned_pointer_is_rejected():
    import ctypes as C
    import torch
    from multigrid_b200 import _cabi
    lib = _cabi.load()
    c = _cabi.MgConfig(8, 8, 2, 7, 100, 0, 0, 148, 0, 1)
    buf = torch.zeros(4096, dtype=torch.int8, device="cuda:0")
    rc = lib.mg_gen_obs(C.byref(c), 1, buf.data_ptr() + 1, buf.data_ptr() + 1024,
                        buf.data_ptr() + 2048, None)
    assert rc == -2
    c.view_size = 4
    assert lib.mg_gen_obs(C.byref(c), 1, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), None) == -1


@pytest.mark.parametrize("path", ["v16", "w32", "w32_unaligned"])
@pytest.mark.parametrize("V,A", [(7, 1000), (3, 5), (9, 33), (5, 4096), (7, 1)])
def test_one_hot_kernel_vs_oracle(V, A, path, monkeypatch):
    """Both one-hot kernels: 16 bytes per thread (16-byte aligned output) and the 4-byte fallback."""
    import ctypes as C
    import torch
    from multigrid_b200 import _cabi
    lib = _cabi.load()
    if path == "w32":
        monkeypatch.setenv("MG_ONE_HOT_W32", "1")
    rng = np.random.default_rng(V * 100 + A)
    stride = _cabi.obs_agent_stride(V)
    img = np.stack([rng.integers(0, 11, (A, V, V)), rng.integers(0, 6, (A, V, V)), rng.integers(0, 4, (A, V, V))], -1)
    buf = np.zeros((A, stride), np.int8)
    buf[:, :3 * V * V] = img.reshape(A, -1)
    obs = torch.from_numpy(buf).cuda()
    raw = torch.full((A * V * V * 21 + 32,), 7, dtype=torch.uint8, device="cuda")
    shift = 4 if path == "w32_unaligned" else 0
    out = raw[shift:shift + A * V * V * 21].view(A, V, V, 21)
    assert lib.mg_one_hot(V, A, stride, obs.data_ptr(), out.data_ptr(), None) == 0
    np.testing.assert_array_equal(out.cpu().numpy(), O.one_hot(img))
    assert (raw[:shift] == 7).all() and (raw[shift + A * V * V * 21:] == 7).all()  # nothing written outside


def test_wrappers_on_gpu():
    import torch
    from multigrid_b200.envs import make
    from multigrid_b200.wrappers import ImgObsWrapper, OneHotObsWrapper, SingleAgentWrapper
    env = make("MultiGrid-Empty-8x8-v0", agents=3, num_envs=50, device="cuda:0")
    obs, _ = env.reset(seed=1)
    plain = {i: obs[i]["image"].cpu().numpy().copy() for i in obs}
    oh = OneHotObsWrapper(env)
    obs, _ = oh.reset(seed=1)
    for i in obs:
        assert obs[i]["image"].shape == (50, 7, 7, 21) and obs[i]["image"].dtype == torch.uint8
        np.testing.assert_array_equal(obs[i]["image"].cpu().numpy(), O.one_hot(plain[i]))
    assert oh.agents[0].observation_space["image"].shape == (7, 7, 21)
    obs, rew, term, trunc, _ = oh.step({0: 2, 1: 1, 2: 0})
    np.testing.assert_array_equal(obs[1]["image"].cpu().numpy(), O.one_hot(env.engine.obs[:, 1].cpu().numpy()))
    img = ImgObsWrapper(env)
    o, _ = img.reset(seed=1)
    assert o[0].shape == (50, 7, 7, 3)
    single = SingleAgentWrapper(make("MultiGrid-Empty-5x5-v0", agents=1, num_envs=8, device="cuda:0"))
    o, info = single.reset(seed=0)
    assert set(o) == {"image", "direction", "mission"}
    o, r, te, tr, info = single.step(2)
    assert r.shape == (8,) and te.dtype == torch.bool


def test_full_obs_kernel_vs_reference_wrapper():
    import ctypes as C
    import torch
    from multigrid_b200 import _cabi
    from tests.hostsim.sim import pack_cells
    lib = _cabi.load()
    d = np.load(f"{GOLDEN_DIR}/full_obs_kat.npz")
    for c in range(len(d["dims"])):
        W, H, n = (int(v) for v in d["dims"][c])
        cells = torch.from_numpy(pack_cells(d["grid"][c:c + 1, :W, :H]).view(np.int32)).cuda()
        agents = torch.from_numpy(O.pack_agents(d["agents"][c:c + 1, :n])).cuda()
        out = torch.zeros((1, W, H, 3), dtype=torch.int8, device="cuda")
        assert lib.mg_full_obs(W, H, n, 1, cells.data_ptr(), agents.data_ptr(), out.data_ptr(), None) == 0
        np.testing.assert_array_equal(out.cpu().numpy()[0], d["img"][c, :W, :H], err_msg=f"state {c}")


def test_fully_obs_wrapper_on_gpu():
    from multigrid_b200.envs import make
    from multigrid_b200.wrappers import FullyObsWrapper
    env = FullyObsWrapper(make("MultiGrid-BlockedUnlockPickup-v0", agents=2, num_envs=40, device="cuda:0",
                               layout_seed=3))
    obs, _ = env.reset(seed=2)
    rng = np.random.default_rng(0)
    for t in range(20):
        obs, *_ = env.step(rng.integers(0, 7, (40, 2)).astype(np.int8))
        base = env.unwrapped
        grid, agents = base.grid.state.cpu().numpy(), base.agent_states.cpu().numpy()
        want = np.stack([O.full_obs(grid[e], agents[e]) for e in range(40)])
        assert obs[0]["image"].shape == (40, 11, 6, 3) and obs[1]["image"] is obs[0]["image"]
        np.testing.assert_array_equal(obs[0]["image"].cpu().numpy(), want)


@pytest.mark.parametrize("size,n", [(5, 2), (6, 3), (8, 8), (16, 12), (9, 30)])
def test_layout_kernel_matches_host_generator(size, n):
    """mg_gen_layouts_empty_random on the GPU vs EmptyLayout.generate driven by numpy generators (which
    tests/test_layouts.py pins to the reference's post-reset states): grids, agents, generator state."""
    import torch
    from multigrid_b200 import layouts as L
    from multigrid_b200.engine import EngineConfig, StepEngine
    from multigrid_b200.env import layout_generator_words
    K = 1500
    mk = lambda: [np.random.default_rng([size, n, k]) for k in range(K)]  # noqa: E731
    gens = mk()
    for g in gens[::3]:
        g.integers(0, 10)  # a buffered 32-bit half in every third generator
    st, inc, buf = layout_generator_words(gens)
    eng = StepEngine(EngineConfig(width=size, height=size, num_agents=n), 4, "cuda:0")
    st2, buf2 = eng.gen_layout_pool_empty_random(st, inc, buf)
    W = size
    grid = eng.pool_grid.view(torch.int8).view(K, W + 1, W + 1, 4)[:, :W, :W, :3].cpu().numpy()
    agents = eng.pool_agents.cpu().numpy()
    layout = L.EmptyLayout(n, size=size, agent_start_pos=None, agent_start_dir=None)
    for k in range(K):
        g, a, _ = layout.generate(gens[k], None)
        np.testing.assert_array_equal(grid[k], g, err_msg=f"layout {k}")
        np.testing.assert_array_equal(agents[k], a, err_msg=f"layout {k}")
    st_h, _, buf_h = layout_generator_words(gens)
    np.testing.assert_array_equal(st2, st_h)
    np.testing.assert_array_equal(buf2, buf_h)
    eng.refresh_layout_pool()  # the NEXT layout of every generator, in place, no host work
    agents2 = eng.pool_agents.cpu().numpy()
    for k in range(0, K, 5):
        g, a, _ = layout.generate(gens[k], None)
        np.testing.assert_array_equal(agents2[k], a, err_msg=f"refreshed layout {k}")


def test_env_reset_device_layouts_equal_host_layouts():
    """'MultiGrid-Empty-Random-6x6-v0': reset() with the pool generated on the device == generated in Python."""
    from multigrid_b200.envs import make
    kw = dict(agents=3, num_envs=2000, device="cuda:0", layout_seed=11, pool_size=2000, allow_agent_overlap=False)
    a = make("MultiGrid-Empty-Random-6x6-v0", **kw)
    b = make("MultiGrid-Empty-Random-6x6-v0", device_layouts=False, **kw)
    assert a.device_layouts and not b.device_layouts
    oa, _ = a.reset(seed=5)
    ob, _ = b.reset(seed=5)
    np.testing.assert_array_equal(a.grid.state.cpu().numpy(), b.grid.state.cpu().numpy())
    np.testing.assert_array_equal(a.agent_states.cpu().numpy(), b.agent_states.cpu().numpy())
    for i in range(3):
        np.testing.assert_array_equal(oa[i]["image"].cpu().numpy(), ob[i]["image"].cpu().numpy())
    assert len({x.tobytes() for x in a.agent_states.cpu().numpy()}) > 1000
    rng = np.random.default_rng(0)
    for t in range(30):
        acts = rng.integers(0, 7, (2000, 3)).astype(np.int8)
        ra, rb = a.step(acts), b.step(acts)
        for i in range(3):
            np.testing.assert_array_equal(ra[0][i]["image"].cpu().numpy(), rb[0][i]["image"].cpu().numpy())
            assert (ra[1][i].cpu().numpy() == rb[1][i].cpu().numpy()).all()


@pytest.mark.parametrize("S,n", [(6, 2), (4, 1), (7, 6)])
def test_bup_layout_kernel_matches_host_generator(S, n):
    """mg_gen_layouts_bup on the GPU vs BlockedUnlockPickupLayout.generate with numpy generators."""
    import torch
    from multigrid_b200 import layouts as L
    from multigrid_b200.engine import EngineConfig, StepEngine
    from multigrid_b200.env import layout_generator_words
    K, W = 1200, 2 * (S - 1) + 1
    lg = [np.random.default_rng([S, n, k]) for k in range(K)]
    og = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(77 * S + k))) for k in range(K)]
    for g in lg[::3]:
        g.integers(0, 10)
    st, inc, buf = layout_generator_words(lg)
    ost, oinc, _ = layout_generator_words(og)
    eng = StepEngine(EngineConfig(width=W, height=S, num_agents=n, hook=1), 4, "cuda:0")
    ost2, info, st2, buf2 = eng.gen_layout_pool_bup(S, st, inc, buf, ost, oinc)
    grid = eng.pool_grid.view(torch.int8).view(K, W + 1, S + 1, 4)[:, :W, :S, :3].cpu().numpy()
    agents = eng.pool_agents.cpu().numpy()
    layout = L.BlockedUnlockPickupLayout(n, room_size=S)
    colors = [c.value for c in L._COLORS]
    for k in range(K):
        g, a, inf = layout.generate(lg[k], og[k])
        np.testing.assert_array_equal(grid[k], g, err_msg=f"layout {k}")
        np.testing.assert_array_equal(agents[k], a, err_msg=f"layout {k}")
        assert inf["mission"] == f"pick up the {colors[info[k]]} box"
    st_h, _, buf_h = layout_generator_words(lg)
    ost_h, _, _ = layout_generator_words(og)
    np.testing.assert_array_equal(st2, st_h)
    np.testing.assert_array_equal(buf2, buf_h)
    np.testing.assert_array_equal(ost2, ost_h)


def test_bup_env_reset_device_layouts_equal_host_layouts():
    from multigrid_b200.envs import make
    kw = dict(agents=2, num_envs=3000, device="cuda:0", layout_seed=3, pool_size=1000, auto_reset=True, max_steps=12)
    a = make("MultiGrid-BlockedUnlockPickup-v0", **kw)
    b = make("MultiGrid-BlockedUnlockPickup-v0", device_layouts=False, **kw)
    a.reset(seed=21)
    b.reset(seed=21)
    np.testing.assert_array_equal(a.grid.state.cpu().numpy(), b.grid.state.cpu().numpy())
    np.testing.assert_array_equal(a.agent_states.cpu().numpy(), b.agent_states.cpu().numpy())
    assert [a.missions[e] for e in range(0, 3000, 37)] == [b.missions[e] for e in range(0, 3000, 37)]
    rng = np.random.default_rng(0)
    for t in range(40):  # through auto-resets from the pool
        acts = rng.integers(0, 7, (3000, 2)).astype(np.int8)
        ra, rb = a.step(acts), b.step(acts)
        for i in range(2):
            np.testing.assert_array_equal(ra[0][i]["image"].cpu().numpy(), rb[0][i]["image"].cpu().numpy())
            assert (ra[1][i].cpu().numpy() == rb[1][i].cpu().numpy()).all()
    np.testing.assert_array_equal(a.grid.state.cpu().numpy(), b.grid.state.cpu().numpy())


def test_rbd_env_reset_device_layouts_equal_host_layouts():
    """'MultiGrid-RedBlueDoors-6x6-v0': pool from mg_gen_layouts_red_blue_doors == pool generated in Python,
    and refresh_layout_pool() continues the same generators."""
    from multigrid_b200 import layouts as L
    from multigrid_b200.envs import make
    kw = dict(agents=3, num_envs=1500, device="cuda:0", layout_seed=8, pool_size=1500)
    a = make("MultiGrid-RedBlueDoors-6x6-v0", **kw)
    b = make("MultiGrid-RedBlueDoors-6x6-v0", device_layouts=False, **kw)
    assert a.device_layouts and not b.device_layouts
    oa, _ = a.reset(seed=2)
    ob, _ = b.reset(seed=2)
    np.testing.assert_array_equal(a.grid.state.cpu().numpy(), b.grid.state.cpu().numpy())
    np.testing.assert_array_equal(a.agent_states.cpu().numpy(), b.agent_states.cpu().numpy())
    np.testing.assert_array_equal(oa[2]["image"].cpu().numpy(), ob[2]["image"].cpu().numpy())
    rng = np.random.default_rng(1)
    for t in range(30):
        acts = rng.integers(0, 7, (1500, 3)).astype(np.int8)
        ra, rb = a.step(acts), b.step(acts)
        np.testing.assert_array_equal(ra[0][0]["image"].cpu().numpy(), rb[0][0]["image"].cpu().numpy())
        assert (ra[1][1].cpu().numpy() == rb[1][1].cpu().numpy()).all()
    a.engine.refresh_layout_pool()  # second layout of generator k == the host generator's second draw
    pool2 = a.engine.pool_agents.cpu().numpy()
    layout = L.RedBlueDoorsLayout(3, size=6)
    for k in range(0, 1500, 50):
        g = np.random.default_rng([8, k])
        layout.generate(g, None)
        _, a2, _ = layout.generate(g, None)
        np.testing.assert_array_equal(pool2[k], a2)


@pytest.mark.parametrize("env_id,n", [("MultiGrid-LockedHallway-6Rooms-v0", 4), ("MultiGrid-LockedHallway-2Rooms-v0", 2),
                                      ("MultiGrid-Playground-v0", 3)])
def test_roomgrid_env_reset_device_layouts_equal_host_layouts(env_id, n):
    """Pools from mg_gen_layouts_locked_hallway / mg_gen_layouts_playground == pools generated in Python
    (which tests/test_layouts.py pins to the reference's post-reset states), then identical rollouts."""
    from multigrid_b200.envs import make
    kw = dict(agents=n, num_envs=600, device="cuda:0", layout_seed=6, pool_size=600)
    a = make(env_id, **kw)
    b = make(env_id, device_layouts=False, **kw)
    assert a.device_layouts and not b.device_layouts
    oa, _ = a.reset(seed=4)
    ob, _ = b.reset(seed=4)
    np.testing.assert_array_equal(a.grid.state.cpu().numpy(), b.grid.state.cpu().numpy())
    np.testing.assert_array_equal(a.agent_states.cpu().numpy(), b.agent_states.cpu().numpy())
    assert len({x.tobytes() for x in a.grid.state.cpu().numpy()}) > 300
    rng = np.random.default_rng(2)
    for t in range(25):
        acts = rng.integers(0, 7, (600, n)).astype(np.int8)
        ra, rb = a.step(acts), b.step(acts)
        for i in range(n):
            np.testing.assert_array_equal(ra[0][i]["image"].cpu().numpy(), rb[0][i]["image"].cpu().numpy())
            assert (ra[1][i].cpu().numpy() == rb[1][i].cpu().numpy()).all()
    np.testing.assert_array_equal(a.agent_states.cpu().numpy(), b.agent_states.cpu().numpy())


def test_reset_where_matches_auto_reset_semantics():
    """mg_reset_where == what the kernel's auto-reset does to an env: next pool layout, counters zeroed,
    PCG stream untouched; unselected envs untouched."""
    import torch
    cfg = O.OracleConfig(W=11, H=6, n=2, V=7, hook=1, auto_reset=True, layout_stride=3, max_steps=50)
    B = 777
    st = random_batch(cfg, B, 12)
    g = GpuEngine(cfg, **st)
    rng = np.random.default_rng(5)
    for t in range(5):
        g.step(rng.integers(0, 7, size=(B, cfg.n)).astype(np.int8))
    before = dict(grid=g.grid, agents=g.agents, sc=g.step_count, pcg=g.pcg_state, idx=g.layout_idx)
    mask = rng.random(B) < 0.3
    g.eng.reset_where(torch.from_numpy(mask).cuda())
    K = st["pool_grid"].shape[0]
    new_idx = np.where(mask, (before["idx"] + 3) % K, before["idx"])
    np.testing.assert_array_equal(g.layout_idx, new_idx)
    np.testing.assert_array_equal(g.step_count, np.where(mask, 0, before["sc"]))
    np.testing.assert_array_equal(g.pcg_state, before["pcg"])
    np.testing.assert_array_equal(g.grid, np.where(mask[:, None, None, None], st["pool_grid"][new_idx], before["grid"]))
    np.testing.assert_array_equal(g.agents, np.where(mask[:, None, None], st["pool_agents"][new_idx], before["agents"]))
    obs = g.gen_obs()  # and the engine keeps stepping consistently with an oracle given the same state
    ora = COracle(cfg, grid=g.grid, agents=g.agents, pcg_state=g.pcg_state, pcg_inc=st["pcg_inc"],
                  pool_grid=st["pool_grid"], pool_agents=st["pool_agents"], layout_idx=g.layout_idx,
                  step_count=g.step_count, nthreads=NTHREADS)
    np.testing.assert_array_equal(obs, ora.gen_obs())
    for t in range(10):
        a = rng.integers(0, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(a)
        o2, r2, t2, tr2 = g.step(a)
        np.testing.assert_array_equal(o2, o1)
        assert (r1 == r2).all()


def _random_config(rng):
    """A random engine configuration: grid 3..24 per side, 1..12 agents, any odd view 3..15, any flag set,
    hooks none / BlockedUnlockPickup / RedBlueDoors / LockedHallway."""
    hook = int(rng.choice([0, 0, 0, 1, 2, 3]))
    kw = dict(W=int(rng.integers(3, 25)), H=int(rng.integers(3, 25)), n=int(rng.integers(1, 13)),
              V=int(rng.choice([3, 5, 7, 9, 11, 13, 15])), max_steps=int(rng.integers(3, 40)),
              see_through_walls=bool(rng.random() < 0.25), allow_agent_overlap=bool(rng.random() < 0.6),
              joint_reward=bool(rng.random() < 0.5), success_any=bool(rng.random() < 0.5),
              failure_any=bool(rng.random() < 0.5), auto_reset=bool(rng.random() < 0.6),
              layout_stride=int(rng.integers(0, 5)), hook=hook)
    if hook == 3:
        kw["hook_param"] = int(rng.integers(1, 5))
    return kw


@pytest.mark.parametrize("case", range(48))
def test_random_configurations_vs_c_oracle(case):
    """Differential sweep over the configuration space (sizes, agent counts, views, flags, hooks):
    the kernel (fused path; every sixth case also through mg_rollout) vs the C oracle on dense random soups."""
    import torch
    rng = np.random.default_rng(10_000 + case)
    kw = _random_config(rng)
    cfg = O.OracleConfig(**kw)
    B, T = int(rng.integers(1, 400)), 14
    st = random_batch(cfg, B, 20_000 + case)
    ora, g = COracle(cfg, nthreads=NTHREADS, **st), GpuEngine(cfg, **st)
    msg = f"case {case}: {kw} B={B}"
    np.testing.assert_array_equal(g.gen_obs(), ora.gen_obs(), err_msg=msg)
    actions = rng.integers(-1, 7, size=(T, B, cfg.n)).astype(np.int8)
    if case % 6 == 5:
        out = g.eng.rollout(torch.from_numpy(actions).cuda())
        g.eng.check_status()
        obs, rew = out["obs"].cpu().numpy(), out["reward"].cpu().numpy()
        term, trunc = out["terminated"].cpu().numpy(), out["truncated"].cpu().numpy()
    for t in range(T):
        o1, r1, t1, tr1 = ora.step(actions[t])
        if case % 6 == 5:
            o2, r2, t2, tr2 = obs[t], rew[t], term[t], trunc[t]
        else:
            o2, r2, t2, tr2 = g.step(actions[t])
        np.testing.assert_array_equal(o2, o1, err_msg=f"{msg} step {t}")
        assert (r1 == r2).all(), f"{msg} step {t}"
        np.testing.assert_array_equal(t2, t1, err_msg=f"{msg} step {t}")
        np.testing.assert_array_equal(tr2, tr1, err_msg=f"{msg} step {t}")
    assert_same(g, ora, msg)


def test_largest_grids():
    """Grids beyond what 16 envs per warp can stage in shared memory fall back to 8 envs per warp (70 x 70
    here); beyond that (127 x 127) the call is refused with MG_ERR_TOO_LARGE, not a launch failure."""
    import ctypes as C
    import torch
    from multigrid_b200 import _cabi
    cfg = O.OracleConfig(W=70, H=70, n=3, V=7, max_steps=30, auto_reset=True)
    B, T = 50, 8
    st = random_batch(cfg, B, 77, K=3)
    ora, g = COracle(cfg, nthreads=NTHREADS, **st), GpuEngine(cfg, **st)
    np.testing.assert_array_equal(g.gen_obs(), ora.gen_obs())
    rng = np.random.default_rng(4)
    for t in range(T):
        a = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(a)
        o2, r2, t2, tr2 = g.step(a)
        np.testing.assert_array_equal(o2, o1)
        assert (r1 == r2).all()
    assert_same(g, ora, "70x70")
    lib = _cabi.load()
    c = _cabi.MgConfig(127, 127, 2, 7, 100, 0, 0, 148, 0, 1)
    buf = torch.zeros(128 * 128 * 4 * 2 + 4096, dtype=torch.int8, device="cuda:0")
    rc = lib.mg_gen_obs(C.byref(c), 2, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), None)
    assert rc == -3 and b"too large" in lib.mg_error_string(rc)


@pytest.mark.parametrize("V,n,E", [(7, 4, 300), (3, 1, 5), (9, 3, 37), (5, 2, 1)])
def test_obs_features_kernel_vs_reference_expression(V, n, E):
    """mg_obs_features == OneHotObsWrapper.one_hot (oracle restatement, pinned to the reference's numba output)
    followed by scripts/train.py:56-63 preprocess_batch restated with torch on the CPU, bit for bit."""
    import torch
    from multigrid_b200.engine import EngineConfig, StepEngine
    cfg = O.OracleConfig(W=9, H=8, n=n, V=V)
    st = random_batch(cfg, E, V * 10 + n)
    g = GpuEngine(cfg, **st)
    img = g.gen_obs()                                   # (E, n, V, V, 3)
    direction = torch.from_numpy(g.agents[..., O.A_DIR].astype(np.int64))
    image = torch.from_numpy(O.one_hot(img))            # uint8 (E, n, V, V, 21)
    d = 2 * torch.pi * (direction / 4)                  # the reference's expression (len(Direction) == 4)
    d = torch.stack([torch.cos(d), torch.sin(d)], dim=-1)
    d = d[..., None, None, :].expand(*image.shape[:-1], 2)
    want = torch.cat([image, d], dim=-1).float()
    got = g.eng.obs_features().cpu()
    assert got.shape == want.shape == (E, n, V, V, 23) and got.dtype == torch.float32
    assert torch.equal(got, want)


@pytest.mark.parametrize("chained", [False, True])
@pytest.mark.parametrize("seed,B,kw", [
    (0, 5000, dict(W=8, H=8, n=4, V=7, auto_reset=True, max_steps=6)),
    (1, 3000, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True, auto_reset=True, max_steps=9)),
    (2, 700, dict(W=9, H=7, n=3, V=5, allow_agent_overlap=False, auto_reset=True, max_steps=5, failure_any=True)),
    (3, 40000, dict(W=8, H=8, n=4, V=7, auto_reset=True, max_steps=7)),
])
def test_single_layout_dedup(seed, B, kw, chained):
    """ONE pool layout (num_layouts == 1): groups whose envs all still equal it load their cells from the
    32-copy buffer instead of their own grids. Random injected grids start dirty; auto-reset makes envs clean;
    pickups / drops / toggles on the (object-rich) pool layout make them dirty again; reset_where and
    load_state flip the flags from the host. Every step vs the C oracle."""
    import torch
    kw = dict(kw)
    cfg = O.OracleConfig(**kw)
    st = random_batch(cfg, B, 300 + seed, K=1)
    ora, g = COracle(cfg, nthreads=NTHREADS, **st), GpuEngine(cfg, **st)
    assert g.eng.pool_rep is not None and bool(g.eng.grid_dirty.all())
    rng = np.random.default_rng(seed)
    T = 30
    for t in range(T):
        a = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(a)
        if chained:
            g.eng.step(torch.from_numpy(a).cuda(), chained=True)
            o2, r2 = g._obs(g.eng.obs_buf), g.eng.reward.cpu().numpy()
        else:
            o2, r2, _, _ = g.step(a)
        np.testing.assert_array_equal(o2, o1, err_msg=f"step {t}")
        assert (r1 == r2).all(), f"step {t}"
        if t == 12:  # host-driven reset of a third of the envs: clean again
            mask = rng.random(B) < 0.33
            g.eng.reset_where(torch.from_numpy(mask).cuda())
            K = 1
            ora.layout_idx[mask] = (ora.layout_idx[mask] + cfg.layout_stride) % K
            ora.grid[mask] = st["pool_grid"][0]
            ora.agents[mask] = st["pool_agents"][0]
            ora.step_count[mask] = 0
            ora.hook_state[mask] = 0
            ora.cell_flags[mask] = 0
            assert not bool(g.eng.grid_dirty[torch.from_numpy(mask).cuda()].any())
    assert_same(g, ora, "dedup")
    dirty = g.eng.grid_dirty.cpu().numpy().astype(bool)
    same = (g.grid == st["pool_grid"][0]).reshape(B, -1).all(1)
    assert same[~dirty].all(), "an env marked clean differs from the pool layout"
    assert (~dirty).sum() > 0, "no env ever became clean: the dedup path was not exercised"
