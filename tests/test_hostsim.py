"""CPU run of the CUDA kernels' phase functions (tests/hostsim) against the reference-recorded
golden fixtures and the C oracle. Catches logic errors before any GPU time is spent; the real
GPU parity tests are in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import mg_oracle as O
from oracle.c_oracle import COracle
from tests.golden_util import ROLLOUT_CASES, load_case, GOLDEN_DIR
from tests.hostsim.sim import SimEngine
from tests.test_oracle_golden import cfg_from_meta
from tests.randstate import random_batch


@pytest.mark.parametrize("variant", ["fused", "generic_view", "group32", "split"])
@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_hostsim_rollout_matches_reference(name, variant):
    d, meta = load_case(name)
    cfg = cfg_from_meta(meta)
    if variant == "split" and (cfg.hook or cfg.auto_reset):
        pytest.skip("split step/gen_obs is only equivalent without post-hook / auto-reset")
    B, T, J = meta["B"], meta["T"], meta["pool_J"]
    T = min(T, 150)
    kw = dict(fused={}, generic_view=dict(generic=1), group32=dict(forced_group=32),
              split=dict(split=True))[variant]
    sim = SimEngine(cfg, d["init_grid"], O.pack_agents(d["init_agents"]), d["pcg_state"],
                    d["pcg_inc"], pool_grid=d["pool_grid"],
                    pool_agents=O.pack_agents(d["pool_agents"]), layout_idx=np.arange(B) * J, **kw)
    np.testing.assert_array_equal(sim.gen_obs(), d["obs0"])
    for t in range(T):
        obs, rew, term, trunc = sim.step(d["actions"][t])
        msg = f"{name} step {t}"
        np.testing.assert_array_equal(sim.grid, d["grid"][t], err_msg=msg)
        np.testing.assert_array_equal(O.unpack_agents(sim.agents), d["agents"][t], err_msg=msg)
        np.testing.assert_array_equal(obs, d["obs"][t], err_msg=msg)
        assert (rew == d["reward"][t]).all(), msg
        np.testing.assert_array_equal(term, d["terminated"][t], err_msg=msg)
        np.testing.assert_array_equal(trunc, d["truncated"][t], err_msg=msg)
        np.testing.assert_array_equal(sim.step_count, d["step_count"][t], err_msg=msg)


def test_hostsim_obs_random_injected_states():
    d = np.load(f"{GOLDEN_DIR}/obs_random.npz")
    for c in range(len(d["W"])):
        W, H, n, V = (int(d[k][c]) for k in ("W", "H", "n", "V"))
        cfg = O.OracleConfig(W=W, H=H, n=n, V=V, see_through_walls=bool(d["stw"][c]))
        grid = np.ascontiguousarray(d["grid"][c, :W, :H])[None]
        agents = O.pack_agents(d["agents"][c, :n])[None]
        z = np.zeros((1, 2), np.uint64)
        for generic in (0, 1):
            got = SimEngine(cfg, grid, agents, z, z, generic=generic).gen_obs()[0]
            np.testing.assert_array_equal(got, d["obs"][c, :n, :V, :V], err_msg=f"case {c}")


@pytest.mark.parametrize("seed,kw", [
    (0, dict(W=8, H=8, n=4, V=7)),
    (1, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True)),
    (2, dict(W=9, H=13, n=5, V=9, allow_agent_overlap=False, failure_any=True)),
    (3, dict(W=5, H=5, n=1, V=3, success_any=False)),
    (4, dict(W=16, H=16, n=8, V=9, joint_reward=True, success_any=False)),
    (5, dict(W=7, H=7, n=3, V=5, see_through_walls=True, auto_reset=True, max_steps=12)),
    (6, dict(W=10, H=6, n=12, V=11, max_steps=30, auto_reset=True, layout_stride=3)),
    (7, dict(W=6, H=9, n=2, V=13, allow_agent_overlap=False)),
    (10, dict(W=12, H=12, n=32, V=15, auto_reset=True, max_steps=9)),
    (11, dict(W=3, H=3, n=1, V=3)),
])
def test_hostsim_random_soup_vs_c_oracle(seed, kw):
    """Bigger ragged batches (tail block, many blocks) of dense random states."""
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    B, T = 203, 45
    st = random_batch(cfg, B, seed)
    ora = COracle(cfg, **st)
    sim = SimEngine(cfg, **st)
    np.testing.assert_array_equal(sim.gen_obs(), ora.gen_obs())
    rng = np.random.default_rng(seed + 100)
    for t in range(T):
        actions = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        o2, r2, t2, tr2 = sim.step(actions)
        msg = f"step {t}"
        np.testing.assert_array_equal(sim.grid, ora.grid, err_msg=msg)
        np.testing.assert_array_equal(sim.agents, ora.agents, err_msg=msg)
        np.testing.assert_array_equal(o2, o1, err_msg=msg)
        assert (r1 == r2).all(), msg
        np.testing.assert_array_equal(t2, t1, err_msg=msg)
        np.testing.assert_array_equal(tr2, tr1, err_msg=msg)
        np.testing.assert_array_equal(sim.step_count, ora.step_count, err_msg=msg)
        np.testing.assert_array_equal(sim.pcg_state, ora.pcg_state, err_msg=msg)
        np.testing.assert_array_equal(sim.layout_idx, ora.layout_idx, err_msg=msg)


@pytest.mark.parametrize("seed,kw", [
    (0, dict(W=8, H=8, n=4, V=7, auto_reset=True, max_steps=9)),
    (1, dict(W=11, H=6, n=2, V=7, hook=1, joint_reward=True, auto_reset=True, max_steps=14)),
    (6, dict(W=10, H=6, n=12, V=11, max_steps=30, auto_reset=True, layout_stride=3)),
    (7, dict(W=6, H=9, n=2, V=13, allow_agent_overlap=False)),
])
def test_hostsim_rollout_equals_single_steps(seed, kw):
    """The kernel's in-launch step loop (mg_rollout): slice t of every output == step t of the
    C oracle, final state == state after T oracle steps."""
    cfg = O.OracleConfig(max_steps=kw.pop("max_steps", 40), **kw)
    B, T = 75, 33
    st = random_batch(cfg, B, seed)
    ora, sim = COracle(cfg, **st), SimEngine(cfg, **st)
    rng = np.random.default_rng(seed + 7)
    actions = rng.integers(-1, 7, size=(T, B, cfg.n)).astype(np.int8)
    obs, dirs, rew, term, trunc = sim.rollout(actions)
    for t in range(T):
        o1, r1, t1, tr1 = ora.step(actions[t])
        msg = f"step {t}"
        np.testing.assert_array_equal(obs[t], o1, err_msg=msg)
        np.testing.assert_array_equal(dirs[t], ora.agents[..., O.A_DIR], err_msg=msg)
        assert (rew[t] == r1).all(), msg
        np.testing.assert_array_equal(term[t], t1, err_msg=msg)
        np.testing.assert_array_equal(trunc[t], tr1, err_msg=msg)
    np.testing.assert_array_equal(sim.grid, ora.grid)
    np.testing.assert_array_equal(sim.agents, ora.agents)
    np.testing.assert_array_equal(sim.step_count, ora.step_count)
    np.testing.assert_array_equal(sim.pcg_state, ora.pcg_state)
    np.testing.assert_array_equal(sim.layout_idx, ora.layout_idx)


@pytest.mark.parametrize("case", range(16))
def test_hostsim_random_configurations_vs_c_oracle(case):
    """The configuration sweep of tests/test_gpu_parity.py::test_random_configurations_vs_c_oracle on the
    CPU build of the kernel's phase functions (smaller batches)."""
    from tests.test_gpu_parity import _random_config
    rng = np.random.default_rng(10_000 + case)
    kw = _random_config(rng)
    cfg = O.OracleConfig(**kw)
    B, T = int(rng.integers(1, 400)) % 60 + 1, 10
    st = random_batch(cfg, B, 20_000 + case)
    ora, sim = COracle(cfg, **st), SimEngine(cfg, **st)
    msg = f"case {case}: {kw} B={B}"
    np.testing.assert_array_equal(sim.gen_obs(), ora.gen_obs(), err_msg=msg)
    for t in range(T):
        actions = rng.integers(-1, 7, size=(B, cfg.n)).astype(np.int8)
        o1, r1, t1, tr1 = ora.step(actions)
        o2, r2, t2, tr2 = sim.step(actions)
        np.testing.assert_array_equal(o2, o1, err_msg=f"{msg} step {t}")
        assert (r1 == r2).all(), f"{msg} step {t}"
        np.testing.assert_array_equal(t2, t1, err_msg=f"{msg} step {t}")
        np.testing.assert_array_equal(tr2, tr1, err_msg=f"{msg} step {t}")
    np.testing.assert_array_equal(sim.grid, ora.grid, err_msg=msg)
    np.testing.assert_array_equal(sim.agents, ora.agents, err_msg=msg)
    np.testing.assert_array_equal(sim.pcg_state, ora.pcg_state, err_msg=msg)
