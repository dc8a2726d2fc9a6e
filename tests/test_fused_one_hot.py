"""The fused one-hot image (MgStepOut.one_hot): the step kernel itself writes
OneHotObsWrapper.one_hot (multigrid/wrappers.py:158-190) of every observation it produces. With MG_TEST_ONE_HOT=1
tests/gpu_adapter.GpuEngine asks for it on every fused step and compares all of it (poisoned beforehand) with the
oracle's restatement applied to the observations -- which the same step compares with the reference fixtures / the C
oracle. Re-runs a cross-section of the parity tests that way: every kernel family (general unrolled and rolled views,
static-grid fast and rolled), ragged batches (spans that do not start or end on 16-byte boundaries), launch knobs.
(tests/hostsim checks the same on the CPU build of the kernel code on every fused step of every hostsim test.)"""
import numpy as np
import pytest

from tests import test_gpu_parity as P
from tests import test_static_path as S
from tests.golden_util import ROLLOUT_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _ask_for_one_hot(monkeypatch):
    monkeypatch.setenv("MG_TEST_ONE_HOT", "1")


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_fused_one_hot_on_reference_fixtures(name):
    P.test_rollout_matches_reference(name, "device")


@pytest.mark.parametrize("seed,B,kw", P.SOUP)
def test_fused_one_hot_on_random_soups(seed, B, kw):
    P.test_random_soup_vs_c_oracle(seed, B, kw)


@pytest.mark.parametrize("B", [1, 15, 16, 17, 31, 32, 33, 129])
def test_fused_one_hot_ragged_batches(B):
    P.test_ragged_batch_sizes(B)


@pytest.mark.parametrize("knobs", [dict(MG_NO_BULK="1"), dict(MG_GROUP="8"), dict(MG_GROUP="32"), dict(MG_GENERIC_VIEW="1"),
                                   dict(MG_WPB="1", MG_NO_BULK="1")],
                         ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()))
@pytest.mark.parametrize("seed,B,kw", [P.SOUP[0], P.SOUP[2], P.SOUP[5]])
def test_fused_one_hot_launch_knobs(seed, B, kw, knobs, monkeypatch):
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    P.test_random_soup_vs_c_oracle(seed, B, kw)


@pytest.mark.parametrize("case", range(0, 48, 2))
def test_fused_one_hot_random_configurations(case):
    if case % 6 == 5:
        pytest.skip("mg_rollout has no one-hot output")
    P.test_random_configurations_vs_c_oracle(case)


@pytest.mark.parametrize("name", S.STATIC_FIXTURES)
def test_fused_one_hot_static_grid_fixtures(name):
    S.test_gpu_static_matches_reference(name, "device")


@pytest.mark.parametrize("seed,B,kw", S.STATIC_RANDOM)
def test_fused_one_hot_static_grid_random(seed, B, kw, monkeypatch):
    S.test_gpu_static_random_vs_c_oracle(seed, B, kw, monkeypatch)


def test_one_hot_wrapper_uses_the_fused_image():
    """OneHotObsWrapper over the base env: no mg_one_hot launch per step, same tensors as the standalone kernel."""
    import torch
    from multigrid_b200 import _cabi
    from multigrid_b200.envs import make
    from multigrid_b200.wrappers import OneHotObsWrapper
    from oracle.mg_oracle import one_hot
    lib = _cabi.load()
    for env_id, n in [("MultiGrid-Empty-8x8-v0", 4), ("MultiGrid-BlockedUnlockPickup-v0", 2)]:
        env = OneHotObsWrapper(make(env_id, agents=n, num_envs=300, device="cuda:0", auto_reset=True))
        obs, _ = env.reset(seed=5)
        eng = env.unwrapped.engine
        np.testing.assert_array_equal(obs[0]["image"].cpu().numpy(), one_hot(eng.obs.cpu().numpy()[:, 0]))
        g = torch.Generator(device="cuda:0").manual_seed(1)
        for t in range(25):
            a = torch.randint(0, 7, (300, n), dtype=torch.int8, device="cuda:0", generator=g)
            before = lib.mg_launch_count()
            obs, rew, term, trunc, _ = env.step(a)
            assert lib.mg_launch_count() - before == 1  # the fused launch and nothing else
            ref = one_hot(eng.obs.cpu().numpy())
            for i in range(n):
                assert obs[i]["image"].shape == (300, 7, 7, 21) and obs[i]["image"].dtype == torch.uint8
                np.testing.assert_array_equal(obs[i]["image"].cpu().numpy(), ref[:, i])
        env.unwrapped.check()
