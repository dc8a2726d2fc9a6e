"""CPU twin of tests/test_fresh_layouts.py: the same env_is_done() / refresh_slot() the CUDA kernel of
mg_refresh_done_layouts calls, run thread by thread on the host (tests/hostsim), must reproduce the multi-episode
rollouts recorded from the unmodified reference -- a NEW _gen_grid draw from the env's own generator at every reset
(base.py:250-301), door positions from the env's order stream including its buffered 32-bit half (roomgrid.py:324)."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

from multigrid_b200 import _cabi
from multigrid_b200.env import layout_generator_words, pcg64_words
from oracle import mg_oracle as O
from tests.golden_util import load_case
from tests.hostsim import sim as S
from tests.test_oracle_golden import cfg_from_meta

# fixture, generator seed of make_golden.run_case, layout family + parameters, state tweaked after the first reset?
CASES = [
    ("empty6r_n3_autoreset", 35, ("empty",), False),
    ("bup_n2_autoreset", 32, ("bup", 6), False),
    ("rbd_n2_autoreset", 42, ("rbd", 6), False),
    ("lh2_n2_autoreset", 53, ("lh", 2, 5, 1, 2), True),
    ("playground_n2_autoreset", 36, ("pg", 7, 3, 3), False),
]
CODES = {"empty": _cabi.LAYOUT_EMPTY_RANDOM, "bup": _cabi.LAYOUT_BUP, "rbd": _cabi.LAYOUT_RED_BLUE_DOORS,
         "lh": _cabi.LAYOUT_LOCKED_HALLWAY, "pg": _cabi.LAYOUT_PLAYGROUND}


@pytest.mark.parametrize("name,seed,family,tweaked", CASES)
def test_hostsim_fresh_layouts_reproduce_the_reference_episodes(name, seed, family, tweaked):
    d, meta = load_case(name)
    cfg = dataclasses.replace(cfg_from_meta(meta), layout_stride=0)  # an env always resets from ITS slot
    B, T, n = meta["B"], meta["T"], meta["n"]
    episodes = 1 + (np.diff(d["step_count"].astype(np.int64), axis=0) < 0).sum(0)
    assert (episodes >= 3).all(), episodes
    lst, linc, lbuf = layout_generator_words([np.random.default_rng(seed * 1000 + b) for b in range(B)])
    ost, oinc = pcg64_words(np.array([seed * 7919 + b for b in range(B)]))  # env.np_random before reset()
    info, obuf = S.aligned((B,), np.int32), None
    if family[0] == "empty":
        g, a, st, buf = S.gen_layouts_empty_random(cfg.W, cfg.H, n, lst, linc, lbuf)
    elif family[0] == "rbd":
        g, a, st, buf = S.gen_layouts_red_blue_doors(family[1], n, lst, linc, lbuf)
    elif family[0] == "lh":
        g, a, st, buf = S.gen_layouts_locked_hallway(*family[1:], n, lst, linc, lbuf)
    elif family[0] == "bup":
        g, a, st, buf, ost, info = S.gen_layouts_bup(family[1], n, lst, linc, lbuf, ost, oinc)
        obuf = S.gen_layouts_bup.order_buf
    else:
        g, a, st, buf, ost = S.gen_layouts_playground(*family[1:], n, lst, linc, lbuf, ost, oinc)
        obuf = S.gen_layouts_playground.order_buf
    np.testing.assert_array_equal(ost, d["pcg_state"])  # the order stream after the first reset's door draws
    if tweaked:  # (the fixture moved keys next to doors after its first reset: state injection)
        g0, a0 = d["init_grid"], O.pack_agents(d["init_agents"])
    else:        # episode 0 is the generated layout itself
        g0, a0 = g, a
        np.testing.assert_array_equal(g, d["init_grid"])
        np.testing.assert_array_equal(np.asarray(a), O.pack_agents(d["init_agents"]))
    eng = S.SimEngine(cfg, g0, a0, d["pcg_state"], d["pcg_inc"], pool_grid=g, pool_agents=a,
                      layout_idx=np.arange(B, dtype=np.int32))
    linc = S.aligned_copy(linc, np.uint64)
    params = list(family[1:]) + [0] * (4 - len(family[1:]))
    gen = _cabi.MgLayoutGen(CODES[family[0]], (C.c_int32 * 4)(*params), S._p(st).value, S._p(linc).value,
                            S._p(buf).value, None if obuf is None else S._p(obuf).value, S._p(info).value)
    np.testing.assert_array_equal(eng.gen_obs(), d["obs0"])
    for t in range(T):
        obs, rew, term, trunc = eng.step(np.ascontiguousarray(d["actions"][t], dtype=np.int8))
        assert S.lib().sim_refresh_done_layouts(C.byref(eng.c), C.c_int64(B), C.byref(eng.state), C.byref(gen)) == 0
        msg = f"{name} step {t}"
        np.testing.assert_array_equal(obs, d["obs"][t], err_msg=msg)
        assert (rew == d["reward"][t]).all(), msg
        np.testing.assert_array_equal(term, d["terminated"][t], err_msg=msg)
        np.testing.assert_array_equal(trunc, d["truncated"][t], err_msg=msg)
        np.testing.assert_array_equal(eng.grid, d["grid"][t], err_msg=msg)
        np.testing.assert_array_equal(O.unpack_agents(eng.agents), d["agents"][t], err_msg=msg)
