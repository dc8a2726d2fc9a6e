"""The reference-facing env layer (multigrid_b200.env / .envs): registry, seeding, dict API.

Rollouts through `make(id, ...)` + `reset(seed)` + `step({agent: action})` are compared with the
fixtures recorded from the unmodified reference (tests/golden/make_golden.py): same env ids, same
kwargs, same generator seeds, same action tapes. CPU runs use the host-simulated kernels
(tests/hostsim) behind the env layer; `-m gpu` runs use the real CUDA engine."""
import numpy as np
import pytest
import torch

import multigrid_b200.env as env_mod
from multigrid_b200.envs import CONFIGURATIONS, NOT_YET, make
from oracle import mg_oracle as O
from tests.golden_util import load_case

REFERENCE_IDS = [  # multigrid/envs/__init__.py:38-52
    'MultiGrid-BlockedUnlockPickup-v0', 'MultiGrid-Empty-5x5-v0', 'MultiGrid-Empty-Random-5x5-v0',
    'MultiGrid-Empty-6x6-v0', 'MultiGrid-Empty-Random-6x6-v0', 'MultiGrid-Empty-8x8-v0',
    'MultiGrid-Empty-16x16-v0', 'MultiGrid-LockedHallway-2Rooms-v0',
    'MultiGrid-LockedHallway-4Rooms-v0', 'MultiGrid-LockedHallway-6Rooms-v0',
    'MultiGrid-Playground-v0', 'MultiGrid-RedBlueDoors-6x6-v0', 'MultiGrid-RedBlueDoors-8x8-v0']

# (fixture, env id, kwargs, generator seed of make_golden.run_case)
CASES = [
    ("empty8_n2", "MultiGrid-Empty-8x8-v0", dict(agents=2), 1),
    ("empty8_n4", "MultiGrid-Empty-8x8-v0", dict(agents=4), 2),
    ("bup_n2", "MultiGrid-BlockedUnlockPickup-v0", dict(agents=2), 3),
    ("empty16_n8_v9", "MultiGrid-Empty-16x16-v0", dict(agents=8, agent_view_size=9), 4),
    ("empty6r_n3_nooverlap_all", "MultiGrid-Empty-Random-6x6-v0",
     dict(agents=3, allow_agent_overlap=False, success_termination_mode="all", agent_start_dir=None), 5),
    ("empty5_n1", "MultiGrid-Empty-5x5-v0", dict(agents=1, agent_view_size=3), 6),
    ("empty8_n4_joint_stw", "MultiGrid-Empty-8x8-v0",
     dict(agents=4, joint_reward=True, see_through_walls=True, agent_view_size=5), 7),
    ("playground_n3", "MultiGrid-Playground-v0", dict(agents=3), 8),
    ("lh6_n4", "MultiGrid-LockedHallway-6Rooms-v0", dict(agents=4), 52),
]
# (fixtures whose initial state was tweaked after reset are covered at the engine level only)


def test_registry_covers_the_reference_ids():
    assert sorted(list(CONFIGURATIONS) + list(NOT_YET)) == sorted(REFERENCE_IDS)
    with pytest.raises(KeyError):
        make("MultiGrid-Nope-v0")


def test_pcg64_words_match_numpy_generators():
    st, inc = env_mod.pcg64_words([0, 7, 123456789])
    for row, seed in enumerate([0, 7, 123456789]):
        ref = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        assert env_mod.generator_words(ref) == (tuple(int(v) for v in st[row]), tuple(int(v) for v in inc[row]))
    st, inc = env_mod.entropy_words(5)
    assert (inc[:, 0] & np.uint64(1)).all()


def test_no_cuda_means_no_env():
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    with pytest.raises(RuntimeError):  # no CPU fallback in the product path
        make("MultiGrid-Empty-8x8-v0", agents=2, num_envs=4)


def rollout_case(name, env_id, kwargs, seed, device, T_max):
    d, meta = load_case(name)
    B, T, n = meta["B"], min(meta["T"], T_max), meta["n"]
    env = make(env_id, num_envs=B, device=device, pool_size=B, **kwargs)
    assert (env.width, env.height, env.max_steps, env.num_agents) == (meta["W"], meta["H"], meta["max_steps"], n)
    obs, infos = env.reset(seed=[seed * 7919 + b for b in range(B)],
                           options=dict(layout_rngs=[np.random.default_rng(seed * 1000 + b) for b in range(B)]))
    np.testing.assert_array_equal(env.grid.state.cpu().numpy(), d["init_grid"])
    np.testing.assert_array_equal(O.unpack_agents(env.agent_states.cpu().numpy()), d["init_agents"])
    for i in range(n):
        np.testing.assert_array_equal(obs[i]["image"].cpu().numpy(), d["obs0"][:, i])
        np.testing.assert_array_equal(obs[i]["direction"].cpu().numpy(), d["dir0"][:, i])
    for t in range(T):
        acts = d["actions"][t]  # (B, n), -1 = id absent from the dict
        if (acts < 0).any():
            actions = torch.from_numpy(acts)
        elif t % 2:
            actions = {i: acts[:, i] for i in range(n)}
        else:
            actions = {i: torch.from_numpy(acts[:, i]) for i in range(n)}
        obs, rew, term, trunc, infos = env.step(actions)
        msg = f"{name} step {t}"
        assert sorted(obs) == sorted(rew) == sorted(term) == sorted(trunc) == list(range(n))
        for i in range(n):
            np.testing.assert_array_equal(obs[i]["image"].cpu().numpy(), d["obs"][t][:, i], err_msg=msg)
            np.testing.assert_array_equal(obs[i]["direction"].cpu().numpy(), d["direction"][t][:, i], err_msg=msg)
            assert (rew[i].cpu().numpy() == d["reward"][t][:, i]).all(), msg
            np.testing.assert_array_equal(term[i].cpu().numpy(), d["terminated"][t][:, i].astype(bool), err_msg=msg)
            np.testing.assert_array_equal(trunc[i].cpu().numpy(), d["truncated"][t].astype(bool), err_msg=msg)
            assert term[i].dtype == torch.bool and rew[i].dtype == torch.float64
    done = env.is_done().cpu().numpy()
    exp = (d["step_count"][T - 1] >= meta["max_steps"]) | d["agents"][T - 1][:, :, 5].astype(bool).all(1)
    np.testing.assert_array_equal(done, exp)
    env.check()
    return env


@pytest.mark.parametrize("name,env_id,kwargs,seed", CASES)
def test_env_api_rollout_hostsim(name, env_id, kwargs, seed, monkeypatch):
    from tests.hostsim.fake_engine import HostSimStepEngine
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    env = rollout_case(name, env_id, kwargs, seed, "cpu", T_max=60)
    a = env.agents[0]
    assert a.observation_space["image"].shape == (a.view_size, a.view_size, 3)
    assert a.action_space.n == 7 and env.unwrapped is env
    assert a.pos.shape == (env.num_envs, 2) and a.carrying.shape == (env.num_envs, 3)
    assert isinstance(env.missions[0], str)


def test_scalar_actions_and_unknown_action(monkeypatch):
    from tests.hostsim.fake_engine import HostSimStepEngine
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    env = make("MultiGrid-Empty-5x5-v0", agents=2, num_envs=3, device="cpu")
    with pytest.raises(RuntimeError):
        env.step({0: 0})
    env.reset(seed=5)
    obs, rew, term, trunc, _ = env.step({0: 2})  # agent 1 absent from the dict: does not act
    assert (env.agents[0].pos.numpy() == (2, 1)).all() and (env.agents[1].pos.numpy() == (1, 1)).all()
    with pytest.raises(ValueError):
        env.step({0: 7})
    assert env.missions[1] == "get to the green goal square"


@pytest.mark.gpu
@pytest.mark.parametrize("name,env_id,kwargs,seed", CASES)
def test_env_api_rollout_gpu(name, env_id, kwargs, seed):
    rollout_case(name, env_id, kwargs, seed, "cuda:0", T_max=10**9)


@pytest.mark.gpu
def test_env_seed_int_matches_per_env_seeds():
    a = make("MultiGrid-Empty-8x8-v0", agents=4, num_envs=64, device="cuda:0")
    b = make("MultiGrid-Empty-8x8-v0", agents=4, num_envs=32, device="cuda:0", first_env=32)
    a.reset(seed=100)
    b.reset(seed=100)  # the second shard of the same global batch
    rng = np.random.default_rng(0)
    for t in range(40):
        acts = torch.from_numpy(rng.integers(0, 7, (64, 4)).astype(np.int8)).cuda()
        oa = a.step(acts)
        ob = b.step(acts[32:].contiguous())
        for i in range(4):
            assert torch.equal(oa[0][i]["image"][32:], ob[0][i]["image"])
            assert torch.equal(oa[1][i][32:], ob[1][i])


def _to_goal(env, agent):
    """Empty-5x5 from (1,1) facing right to the goal at (3,3): forward x2, right, forward x2."""
    out = None
    for a in (2, 2, 1, 2, 2):
        out = env.step({agent: a})
    return out


def test_rllib_adapter_surface(monkeypatch):
    """multigrid/rllib/__init__.py:44-105: agents / possible_agents, '__all__' = all() over agents, spaces."""
    from tests.hostsim.fake_engine import HostSimStepEngine
    from multigrid_b200.rllib import RLlibWrapper, to_rllib_env
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    cls = to_rllib_env("MultiGrid-Empty-5x5-v0", default_config=dict(agents=2, success_termination_mode="all"))
    assert cls.__name__ == "RLlib_MultiGrid-Empty-5x5-v0"
    env = cls(dict(num_envs=3, device="cpu", max_steps=9))
    assert isinstance(env, RLlibWrapper) and env.agents == [0, 1] == env.possible_agents
    assert env.get_observation_space(1)["image"].shape == (7, 7, 3) and env.get_action_space(0).n == 7
    obs, infos = env.reset(seed=3)
    assert sorted(obs) == [0, 1]
    obs, rew, term, trunc, infos = _to_goal(env, 0)
    assert term[0].all() and not term[1].any() and not term["__all__"].any() and term["__all__"].shape == (3,)
    assert not trunc["__all__"].any() and (rew[0] > 0).all() and (rew[1] == 0).all()
    obs, rew, term, trunc, infos = _to_goal(env, 1)  # step 10 > max_steps 9
    assert term["__all__"].all() and trunc["__all__"].all() and sorted(obs) == [0, 1]


def test_pettingzoo_adapter_surface(monkeypatch):
    """multigrid/pettingzoo/__init__.py:38-115: live-agent list, possible_agents, spaces by id."""
    from tests.hostsim.fake_engine import HostSimStepEngine
    from multigrid_b200.pettingzoo import PettingZooWrapper, to_pettingzoo_env
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    cls = to_pettingzoo_env("MultiGrid-Empty-5x5-v0", metadata={"name": "empty_v0"})
    env = cls(agents=2, num_envs=4, device="cpu", success_termination_mode="all", max_steps=50)
    assert isinstance(env, PettingZooWrapper) and cls.metadata == {"name": "empty_v0"}
    assert env.possible_agents == [0, 1] and sorted(env.observation_spaces) == [0, 1]
    assert env.action_space(1).n == 7 and env.observation_space(0)["direction"].n == 4
    env.reset(seed=0)
    assert env.agents == [0, 1] and env.agent_mask.all()
    _to_goal(env, 0)
    assert env.agents == [1] and not env.agent_mask[:, 0].any() and env.agent_mask[:, 1].all()
    _to_goal(env, 1)
    assert env.agents == [] and not env.agent_mask.any()  # every env is done
    assert env.render_mode is None


def test_wrappers_compose_with_adapters(monkeypatch):
    """ImgObs / SingleAgent wrappers under the adapters (the RLlib registration wraps with a
    wrapper before adapting, rllib/__init__.py:110-111)."""
    from tests.hostsim.fake_engine import HostSimStepEngine
    from multigrid_b200.rllib import to_rllib_env
    from multigrid_b200.wrappers import ImgObsWrapper
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    env = to_rllib_env("MultiGrid-Empty-6x6-v0", ImgObsWrapper)(dict(agents=3, num_envs=2, device="cpu"))
    obs, _ = env.reset(seed=1)
    assert obs[2].shape == (2, 7, 7, 3)
    obs, rew, term, trunc, _ = env.step({0: 2, 1: 0, 2: 1})
    assert obs[0].shape == (2, 7, 7, 3) and "__all__" in term and "__all__" in trunc


@pytest.mark.gpu
def test_adapters_on_gpu():
    """The RLlib registration path of the reference (OneHotObsWrapper under the adapter) and the
    PettingZoo adapter over the real CUDA engine."""
    from multigrid_b200.pettingzoo import to_pettingzoo_env
    from multigrid_b200.rllib import to_rllib_env
    from multigrid_b200.wrappers import OneHotObsWrapper
    env = to_rllib_env("MultiGrid-Empty-5x5-v0", OneHotObsWrapper,
                       default_config=dict(agents=2, success_termination_mode="all"))(
        dict(num_envs=100, device="cuda:0"))
    obs, _ = env.reset(seed=0)
    assert obs[1]["image"].shape == (100, 7, 7, 21) and obs[1]["image"].dtype == torch.uint8
    assert env.get_observation_space(0)["image"].shape == (7, 7, 21)
    obs, rew, term, trunc, _ = _to_goal(env, 0)
    assert term[0].all() and not term["__all__"].any()
    plain = env.env.unwrapped.engine.obs[:, 1].cpu().numpy()
    np.testing.assert_array_equal(obs[1]["image"].cpu().numpy(), O.one_hot(plain))
    pz = to_pettingzoo_env("MultiGrid-Empty-5x5-v0")(agents=2, num_envs=64, device="cuda:0",
                                                     success_termination_mode="all")
    pz.reset(seed=1)
    _to_goal(pz, 1)
    assert pz.agents == [0] and pz.possible_agents == [0, 1]


def test_vectorised_seeding_matches_numpy():
    """seed_sequence_pcg64_words == PCG64(SeedSequence(...)) for int seeds (1 and 2 words) and list entropy."""
    seeds = [0, 1, 7, 123456789, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 5, 2 ** 62 + 11]
    st, inc = env_mod.pcg64_words(seeds)
    for r, s in enumerate(seeds):
        ref = np.random.Generator(np.random.PCG64(np.random.SeedSequence(s)))
        assert env_mod.generator_words(ref) == (tuple(int(v) for v in st[r]), tuple(int(v) for v in inc[r]))
    ent = np.array([[5, 3], [11, 4096], [0, 0], [2 ** 32 - 1, 17]], np.uint32)
    st, inc = env_mod.seed_sequence_pcg64_words(ent)
    for r in range(len(ent)):
        ref = np.random.default_rng([int(ent[r, 0]), int(ent[r, 1])])
        assert env_mod.generator_words(ref) == (tuple(int(v) for v in st[r]), tuple(int(v) for v in inc[r]))
    ent = (np.arange(12, dtype=np.uint32).reshape(2, 6) + 1) * np.uint32(2654435761)
    st, inc = env_mod.seed_sequence_pcg64_words(ent)
    for r in range(2):
        ref = np.random.default_rng([int(x) for x in ent[r]])
        assert env_mod.generator_words(ref) == (tuple(int(v) for v in st[r]), tuple(int(v) for v in inc[r]))


def test_reset_device_layout_path_equals_host_path(monkeypatch):
    """Empty-Random reset with the pool from the layout kernel's function (hostsim) and vectorised seeding
    == the Python generator path with one numpy Generator per layout."""
    from tests.hostsim.fake_engine import HostSimStepEngine
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    kw = dict(agents=3, num_envs=300, device="cpu", layout_seed=11, pool_size=300, first_env=1000)
    a = make("MultiGrid-Empty-Random-5x5-v0", **kw)
    b = make("MultiGrid-Empty-Random-5x5-v0", device_layouts=False, **kw)
    assert a.device_layouts and not b.device_layouts
    oa, _ = a.reset(seed=5)
    ob, _ = b.reset(seed=5)
    np.testing.assert_array_equal(a.grid.state.numpy(), b.grid.state.numpy())
    np.testing.assert_array_equal(a.agent_states.numpy(), b.agent_states.numpy())
    np.testing.assert_array_equal(oa[2]["image"].numpy(), ob[2]["image"].numpy())
    assert len({x.tobytes() for x in a.agent_states.numpy()}) > 100


def test_bup_reset_device_layout_path_equals_host_path(monkeypatch):
    """BlockedUnlockPickup reset with the pool from the BUP layout function (hostsim) == the Python
    generator path: grids, agents, missions, per-env order streams (advanced by the door-height draw)."""
    from tests.hostsim.fake_engine import HostSimStepEngine
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    kw = dict(agents=2, num_envs=60, device="cpu", layout_seed=4, pool_size=40)
    a = make("MultiGrid-BlockedUnlockPickup-v0", **kw)
    b = make("MultiGrid-BlockedUnlockPickup-v0", device_layouts=False, **kw)
    assert a.device_layouts and not b.device_layouts
    oa, _ = a.reset(seed=9)
    ob, _ = b.reset(seed=9)
    np.testing.assert_array_equal(a.grid.state.numpy(), b.grid.state.numpy())
    np.testing.assert_array_equal(a.agent_states.numpy(), b.agent_states.numpy())
    np.testing.assert_array_equal(oa[1]["image"].numpy(), ob[1]["image"].numpy())
    assert [a.missions[e] for e in range(60)] == [b.missions[e] for e in range(60)]
    assert len({a.missions[e] for e in range(60)}) > 1
    rng = np.random.default_rng(0)
    for t in range(25):  # the order streams must agree too: same agent order every step
        acts = rng.integers(0, 7, (60, 2)).astype(np.int8)
        ra, rb = a.step(acts), b.step(acts)
        np.testing.assert_array_equal(ra[0][0]["image"].numpy(), rb[0][0]["image"].numpy())
        np.testing.assert_array_equal(a.agent_states.numpy(), b.agent_states.numpy())


def test_make_accepts_an_iterable_of_agents(monkeypatch):
    """gym.make(id, agents=Iterable[Agent]) (base.py:85-103): the iterable's length is the agent count."""
    from tests.hostsim.fake_engine import HostSimStepEngine
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)

    class FakeAgent:
        def __init__(self, view_size=7):
            self.view_size = view_size

    env = make("MultiGrid-Empty-5x5-v0", agents=[FakeAgent(), FakeAgent(), FakeAgent()], num_envs=2, device="cpu")
    assert env.num_agents == 3 and len(env.agents) == 3
    with pytest.raises(ValueError):
        make("MultiGrid-Empty-5x5-v0", agents=[FakeAgent(5)], num_envs=2, device="cpu")
    with pytest.raises(TypeError):
        make("MultiGrid-Empty-5x5-v0", agents=2.5, num_envs=2, device="cpu")


@pytest.mark.gpu
def test_wide_action_tensors_do_not_wrap_into_valid_actions():
    """int64 actions outside -1..6 (256 would narrow to 0 = left) are flagged like the reference's ValueError."""
    env = make("MultiGrid-Empty-5x5-v0", agents=2, num_envs=8, device="cuda:0")
    env.reset(seed=3)
    before = env.agent_states.clone()
    env.step(torch.full((8, 2), 256, dtype=torch.int64, device="cuda:0"))
    with pytest.raises(ValueError):
        env.check()
    assert torch.equal(env.agent_states, before)  # nobody turned left
    env.step(np.full((8, 2), -3, dtype=np.int64))
    with pytest.raises(ValueError):
        env.check()
    with pytest.raises(TypeError):
        env.engine.step(torch.zeros((4, 2), dtype=torch.int8, device="cuda:0"))  # wrong batch size


@pytest.mark.gpu
def test_one_hot_over_fully_obs_encodes_the_wrapped_image():
    """OneHotObsWrapper(FullyObsWrapper(env)): the one-hot of the whole-grid image, not of the partial views
    (wrappers.py:176-177 encodes whatever image the wrapped env produced)."""
    from multigrid_b200.wrappers import FullyObsWrapper, OneHotObsWrapper
    env = OneHotObsWrapper(FullyObsWrapper(make("MultiGrid-BlockedUnlockPickup-v0", agents=2, num_envs=33,
                                                device="cuda:0")))
    obs, _ = env.reset(seed=4)
    base = env.unwrapped
    for t in range(5):
        obs, *_ = env.step({0: 2, 1: t % 3}, chained=False)
        full = np.stack([O.full_obs(g, a) for g, a in zip(base.grid.state.cpu().numpy(),
                                                          base.agent_states.cpu().numpy())])
        for i in (0, 1):
            got = obs[i]["image"].cpu().numpy()
            assert got.shape == (33, base.width, base.height, 21)
            np.testing.assert_array_equal(got, np.stack([O.one_hot(f) for f in full]), err_msg=f"step {t}")


def test_world_object_view_matches_the_reference_encodings():
    """core/world_object.py: the reference's class names, encodings and rule predicates (world_object.py:28-605)."""
    from multigrid_b200.core.objects import Ball, Box, Door, Floor, Goal, Key, Lava, Wall, WorldObj
    assert Wall().encode() == (2, 5, 0) and Goal().encode() == (8, 1, 0) and Lava().encode() == (9, 0, 0)
    assert Floor("purple").encode() == (3, 3, 0) and Key("yellow").encode() == (5, 4, 0)
    assert Ball("green").encode() == (6, 1, 0) and Box("red").encode() == (7, 0, 0)
    assert Door("red").encode() == (4, 0, 1) and Door("blue", is_open=True).encode() == (4, 2, 0)
    assert Door("grey", is_locked=True).encode() == (4, 5, 2)
    # can_overlap: empty, floor, goal, lava, OPEN door; can_pickup: key, ball, box
    assert [o.can_overlap() for o in (WorldObj.empty(), Floor(), Goal(), Lava(), Door(is_open=True), Door(), Wall(), Key())] == \
        [True, True, True, True, True, False, False, False]
    assert [o.can_pickup() for o in (Key(), Ball(), Box(), Door(), Wall(), Goal())] == [True, True, True, False, False, False]
    assert WorldObj.from_array((1, 0, 0)) is None
    d = WorldObj.from_array((4, 3, 2))
    assert isinstance(d, Door) and d.is_locked and not d.is_open and d.color.value == "purple"
    d.is_locked = False
    d.is_open = True
    assert d.encode() == (4, 3, 0)
    with pytest.raises(ValueError):
        WorldObj.from_array((11, 0, 0))
    assert Key() != Key()  # identity, like the reference


def test_grid_get_set_over_the_tensor_state(monkeypatch):
    """env.grid.get / set (core/grid.py:102-131) read and write single cells of the batched tensor state."""
    from multigrid_b200.core.objects import Ball, Door, Goal, Wall
    from tests.hostsim.fake_engine import HostSimStepEngine
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    env = make("MultiGrid-Empty-8x8-v0", agents=2, num_envs=3, device="cpu")
    env.reset(seed=0)
    assert isinstance(env.grid.get(0, 0, 0), Wall) and isinstance(env.grid.get(2, 6, 6), Goal)
    assert env.grid.get(1, 3, 3) is None and env.grid.get(1, -1, 3) is None
    env.grid.set(1, 3, 3, Ball("purple"))
    env.grid.set(2, 2, 5, Door("red", is_locked=True))
    env.grid.set(0, 6, 6, None)
    assert env.grid.get(1, 3, 3).encode() == (6, 3, 0) and env.grid.get(0, 3, 3) is None
    assert env.grid.get(2, 2, 5).is_locked and env.grid.get(0, 6, 6) is None
    assert tuple(int(v) for v in env.grid.state[1, 3, 3]) == (6, 3, 0)
