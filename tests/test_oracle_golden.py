"""Pin the CPU oracle (oracle/mg_oracle.py) to fixtures recorded from the unmodified reference."""
import numpy as np
import pytest

from oracle import mg_oracle as O
from oracle.c_oracle import COracle
from tests.golden_util import ROLLOUT_CASES, load_case, GOLDEN_DIR


def cfg_from_meta(meta):
    return O.OracleConfig(
        W=meta["W"], H=meta["H"], n=meta["n"], V=meta["V"], max_steps=meta["max_steps"],
        see_through_walls=bool(meta["see_through_walls"]),
        allow_agent_overlap=bool(meta["allow_agent_overlap"]),
        joint_reward=bool(meta["joint_reward"]), success_any=bool(meta["success_any"]),
        failure_any=bool(meta["failure_any"]), hook=int(meta["hook"]),
        hook_param=int(meta.get("hook_param", 0)),
        auto_reset=bool(meta["auto_reset"]), layout_stride=1)


def test_pcg64_known_answers():
    d = np.load(f"{GOLDEN_DIR}/pcg64_kat.npz")
    for i in range(len(d["seeds"])):
        s = int(d["state"][i, 0]) | (int(d["state"][i, 1]) << 64)
        inc = int(d["inc"][i, 0]) | (int(d["inc"][i, 1]) << 64)
        for want in d["draws"][i]:
            s, u = O.pcg64_next_double(s, inc)
            assert u == want


@pytest.mark.parametrize("impl", ["python", "c"])
def test_obs_random_injected_states(impl):
    d = np.load(f"{GOLDEN_DIR}/obs_random.npz")
    for c in range(len(d["W"])):
        W, H, n, V = (int(d[k][c]) for k in ("W", "H", "n", "V"))
        cfg = O.OracleConfig(W=W, H=H, n=n, V=V, see_through_walls=bool(d["stw"][c]))
        grid = np.ascontiguousarray(d["grid"][c, :W, :H])
        agents = O.pack_agents(d["agents"][c, :n])
        if impl == "python":
            got = O.gen_obs_env(cfg, grid, agents)
        else:
            z = np.zeros((1, 2), np.uint64)
            got = COracle(cfg, grid[None], agents[None], z, z).gen_obs()[0]
        np.testing.assert_array_equal(got, d["obs"][c, :n, :V, :V], err_msg=f"case {c}")


@pytest.mark.parametrize("impl", ["python", "c"])
@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_rollout_matches_reference(name, impl):
    d, meta = load_case(name)
    cfg = cfg_from_meta(meta)
    B, T, J = meta["B"], meta["T"], meta["pool_J"]
    cls = O.OracleBatch if impl == "python" else COracle
    ob = cls(cfg, d["init_grid"], O.pack_agents(d["init_agents"]), d["pcg_state"],
                       d["pcg_inc"], pool_grid=d["pool_grid"],
                       pool_agents=O.pack_agents(d["pool_agents"]),
                       layout_idx=np.arange(B) * J)
    np.testing.assert_array_equal(ob.gen_obs(), d["obs0"])
    for t in range(T):
        obs, rew, term, trunc = ob.step(d["actions"][t])
        msg = f"{name} step {t}"
        np.testing.assert_array_equal(ob.grid, d["grid"][t], err_msg=msg)
        np.testing.assert_array_equal(O.unpack_agents(ob.agents), d["agents"][t], err_msg=msg)
        np.testing.assert_array_equal(obs, d["obs"][t], err_msg=msg)
        np.testing.assert_array_equal(ob.agents[..., O.A_DIR], d["direction"][t], err_msg=msg)
        assert (rew == d["reward"][t]).all(), msg  # bit-exact float64
        np.testing.assert_array_equal(term, d["terminated"][t], err_msg=msg)
        np.testing.assert_array_equal(trunc, d["truncated"][t], err_msg=msg)
        np.testing.assert_array_equal(ob.step_count, d["step_count"][t], err_msg=msg)


def test_one_hot_matches_reference_numba():
    d = np.load(f"{GOLDEN_DIR}/one_hot_kat.npz")
    for c in range(len(d["V"])):
        V = int(d["V"][c])
        np.testing.assert_array_equal(O.one_hot(d["x"][c, :V, :V]), d["out"][c, :V, :V])


def test_full_obs_matches_reference_wrapper():
    d = np.load(f"{GOLDEN_DIR}/full_obs_kat.npz")
    for c in range(len(d["dims"])):
        W, H, n = (int(v) for v in d["dims"][c])
        got = O.full_obs(d["grid"][c, :W, :H], O.pack_agents(d["agents"][c, :n]))
        np.testing.assert_array_equal(got, d["img"][c, :W, :H], err_msg=f"state {c}")
