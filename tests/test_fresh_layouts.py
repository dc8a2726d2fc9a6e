"""Fresh layouts on auto-reset (mg_refresh_done_layouts, StepEngine.enable_fresh_layouts): every episode of every env
gets a NEW _gen_grid draw from the env's own generator, as the reference's reset() does (base.py:250-301). Checked
against multi-episode rollouts recorded from the unmodified reference (tests/golden/*_autoreset.npz: each env's
RandomMixin generator is default_rng(seed*1000 + b) and keeps drawing across its resets): the engine is given only
those generators and must reproduce every later episode's layout -- grids, agent placements, observations -- itself."""
import numpy as np
import pytest

from oracle import mg_oracle as O
from tests.golden_util import load_case
from tests.test_oracle_golden import cfg_from_meta

pytestmark = pytest.mark.gpu

# fixture, generator seed of make_golden.run_case, layout family + parameters, state tweaked after the first reset?
CASES = [
    ("empty6r_n3_autoreset", 35, ("empty",), False),
    ("bup_n2_autoreset", 32, ("bup", 6), False),
    ("rbd_n2_autoreset", 42, ("rbd", 6), False),
    ("lh2_n2_autoreset", 53, ("lh", 2, 5, 1, 2), True),
    ("playground_n2_autoreset", 36, ("pg", 7, 3, 3), False),
]


@pytest.mark.parametrize("name,seed,family,tweaked", CASES)
def test_fresh_layouts_reproduce_the_reference_episodes(name, seed, family, tweaked):
    import torch
    from multigrid_b200.engine import StepEngine
    from multigrid_b200.env import layout_generator_words, pcg64_words
    from tests.gpu_adapter import engine_config
    d, meta = load_case(name)
    cfg = cfg_from_meta(meta)
    B, T = meta["B"], meta["T"]
    episodes = 1 + (np.diff(d["step_count"].astype(np.int64), axis=0) < 0).sum(0)
    assert (episodes >= 3).all(), episodes  # several episodes per env
    eng = StepEngine(engine_config(cfg), B, "cuda:0")
    lst, linc, lbuf = layout_generator_words([np.random.default_rng(seed * 1000 + b) for b in range(B)])
    ost, oinc = pcg64_words(np.array([seed * 7919 + b for b in range(B)]))  # env.np_random before reset()
    if family[0] == "empty":
        eng.gen_layout_pool_empty_random(lst, linc, lbuf)
    elif family[0] == "rbd":
        eng.gen_layout_pool_red_blue_doors(family[1], lst, linc, lbuf)
    elif family[0] == "lh":
        eng.gen_layout_pool_locked_hallway(*family[1:], lst, linc, lbuf)
    elif family[0] == "bup":
        ost = eng.gen_layout_pool_bup(family[1], lst, linc, lbuf, ost, oinc)[0]
    else:
        ost = eng.gen_layout_pool_playground(*family[1:], lst, linc, lbuf, ost, oinc)[0]
    np.testing.assert_array_equal(ost, d["pcg_state"])  # the order stream after the first reset's door draws
    eng.load_state(layout_idx=np.arange(B, dtype=np.int32), pcg_state=d["pcg_state"], pcg_inc=d["pcg_inc"])
    eng.reset_from_pool()
    if not tweaked:  # episode 0 is the device-generated layout itself
        np.testing.assert_array_equal(eng.grid.cpu().numpy(), d["init_grid"])
        np.testing.assert_array_equal(eng.agents.cpu().numpy(), O.pack_agents(d["init_agents"]))
    else:            # (the fixture moved keys next to doors after its first reset: state injection)
        eng.load_state(grid=d["init_grid"], agents=O.pack_agents(d["init_agents"]))
    eng.enable_fresh_layouts()
    V = cfg.V
    np.testing.assert_array_equal(eng.gen_obs().cpu().numpy(), d["obs0"])
    for t in range(T):
        obs, rew, term, trunc = eng.step(torch.from_numpy(np.ascontiguousarray(d["actions"][t], dtype=np.int8)).cuda())
        msg = f"{name} step {t}"
        np.testing.assert_array_equal(obs.cpu().numpy(), d["obs"][t], err_msg=msg)
        assert (rew.cpu().numpy() == d["reward"][t]).all(), msg
        np.testing.assert_array_equal(term.cpu().numpy(), d["terminated"][t], err_msg=msg)
        np.testing.assert_array_equal(trunc.cpu().numpy(), d["truncated"][t], err_msg=msg)
        np.testing.assert_array_equal(eng.grid.cpu().numpy(), d["grid"][t], err_msg=msg)
        np.testing.assert_array_equal(O.unpack_agents(eng.agents.cpu().numpy()), d["agents"][t], err_msg=msg)
    eng.check_status()


def test_env_level_fresh_layouts_differ_between_episodes():
    """make(..., auto_reset=True, fresh_layouts=True): consecutive episodes of an env start from different layouts
    (a cycling pool of one slot per env would repeat the same one)."""
    import torch
    from multigrid_b200.envs import make
    env = make("MultiGrid-Empty-Random-6x6-v0", agents=2, num_envs=64, device="cuda:0", auto_reset=True,
               fresh_layouts=True, max_steps=5, layout_seed=3)
    env.reset(seed=0)
    starts = []
    for t in range(6 * 4):
        env.step(torch.full((64, 2), 6, dtype=torch.int8, device="cuda:0"))  # `done`: nobody moves
        if env.step_count[0].item() == 0:
            starts.append(env.agent_states[:, :, :3].cpu().numpy().copy())
    assert len(starts) >= 3
    assert not np.array_equal(starts[0], starts[1]) and not np.array_equal(starts[1], starts[2])
    env.check()
