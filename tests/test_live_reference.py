"""Randomised differential test against the LIVE reference: fresh rollouts of the unmodified reference
(/root/reference through tests/golden/shims, exactly how the committed fixtures were recorded) with a NEW random
seed every run, replayed on the C oracle and on the CPU build of the CUDA kernels' phase functions (tests/hostsim).
Skipped where the reference tree does not exist (the GPU box); set MG_LIVE_SEED to reproduce a failing run."""
import os
import sys

import numpy as np
import pytest

REFERENCE = os.environ.get("MULTIGRID_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "multigrid")),
                                reason="the reference tree is not present (GPU box)")

SEED = int(os.environ.get("MG_LIVE_SEED", str(int.from_bytes(os.urandom(3), "little"))))

CASES = [
    ("MultiGrid-Empty-8x8-v0", dict(agents=4), dict()),
    ("MultiGrid-Empty-Random-6x6-v0", dict(agents=3, allow_agent_overlap=False, success_termination_mode="all"),
     dict(p_absent=0.1)),
    ("MultiGrid-BlockedUnlockPickup-v0", dict(agents=2), dict(fwd_heavy=True)),
    ("MultiGrid-RedBlueDoors-6x6-v0", dict(agents=3), dict(fwd_heavy=True)),
    ("MultiGrid-LockedHallway-4Rooms-v0", dict(agents=2, max_steps=30), dict(auto_reset=True, fwd_heavy=True)),
    ("MultiGrid-Playground-v0", dict(agents=2, allow_agent_overlap=False), dict(fwd_heavy=True)),
    ("Golden-Soup-v0", dict(agents=5, joint_reward=True, agent_view_size=5, max_steps=25),
     dict(auto_reset=True, p_absent=0.05)),
    ("MultiGrid-Empty-16x16-v0", dict(agents=8, agent_view_size=9, see_through_walls=True), dict()),
]


@pytest.fixture(scope="module")
def golden():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden  # executes the reference's imports (numba JIT on first use)
    return make_golden


@pytest.mark.parametrize("case", range(len(CASES)))
def test_fresh_reference_rollout(case, golden):
    from oracle import mg_oracle as O
    from oracle.c_oracle import COracle
    from tests.hostsim.sim import SimEngine
    from tests.test_oracle_golden import cfg_from_meta
    env_id, kwargs, opts = CASES[case]
    seed = SEED + 101 * case
    action_p = golden.FWD_HEAVY if opts.get("fwd_heavy") else None
    rec, meta = golden.run_case(f"live_{case}", env_id, kwargs, B=3, T=60, seed=seed, p_absent=opts.get("p_absent", 0.0),
                                action_p=action_p, auto_reset=opts.get("auto_reset", False), save=False)
    cfg = cfg_from_meta(meta)
    B, T, J = meta["B"], meta["T"], meta["pool_J"]
    engines = [("c-oracle", COracle, {}), ("hostsim", SimEngine, {})]
    pg, pa = rec["pool_grid"], O.pack_agents(rec["pool_agents"])
    if env_id.startswith("MultiGrid-Empty-") and "Random" not in env_id and (pg == pg[:1]).all() and (pa == pa[:1]).all():
        # a grid no action can change: also through the static-grid kernels' code (memoised views, order-independent
        # fast path with the PCG64 jump-ahead, serial path on goal steps)
        engines.append(("hostsim-static", SimEngine, dict(static=True)))
    for name, cls, extra in engines:
        if extra:
            ob = cls(cfg, rec["init_grid"], O.pack_agents(rec["init_agents"]), rec["pcg_state"], rec["pcg_inc"],
                     pool_grid=pg[:1], pool_agents=pa[:1], layout_idx=np.zeros(B, np.int32), **extra)
        else:
            ob = cls(cfg, rec["init_grid"], O.pack_agents(rec["init_agents"]), rec["pcg_state"], rec["pcg_inc"],
                     pool_grid=pg, pool_agents=pa, layout_idx=np.arange(B) * J)
        msg0 = f"{name} {env_id} {kwargs} MG_LIVE_SEED={SEED}"
        np.testing.assert_array_equal(ob.gen_obs(), rec["obs0"], err_msg=msg0)
        for t in range(T):
            obs, rew, term, trunc = ob.step(rec["actions"][t])
            msg = f"{msg0} step {t}"
            np.testing.assert_array_equal(obs, rec["obs"][t], err_msg=msg)
            assert (rew == rec["reward"][t]).all(), msg  # bit-exact float64
            np.testing.assert_array_equal(term, rec["terminated"][t], err_msg=msg)
            np.testing.assert_array_equal(trunc, rec["truncated"][t], err_msg=msg)
            np.testing.assert_array_equal(ob.grid, rec["grid"][t], err_msg=msg)
            np.testing.assert_array_equal(O.unpack_agents(ob.agents), rec["agents"][t], err_msg=msg)
