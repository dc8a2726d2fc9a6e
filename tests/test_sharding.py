"""N>1 path on CPU: two `gloo` ranks each own a shard of one global env batch (host-simulated
kernels behind the env layer). Checks: shard bounds, per-env seeds are a function of the global
env id (shard results == slices of the 1-rank results), max/sum aggregation over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multigrid_b200 import sharding


def test_shard_bounds_partition():
    for total, world in [(65536, 8), (10, 3), (7, 8), (524288, 8)]:
        spans = [sharding.shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rollout(env, tape):
    out = []
    for t in range(tape.shape[0]):
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(tape[t]))
        out.append((np.stack([obs[i]["image"].numpy() for i in obs], 1).copy(),
                    np.stack([rew[i].numpy() for i in rew], 1).copy(),
                    np.stack([term[i].numpy() for i in term], 1).copy()))
    return out


def _worker(rank, world_size, port, total, tape, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import multigrid_b200.env as env_mod
    from tests.hostsim.fake_engine import HostSimStepEngine
    env_mod.StepEngine = HostSimStepEngine
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        env = sharding.make_sharded("MultiGrid-Empty-8x8-v0", total, agents=4, device="cpu",
                                    auto_reset=True, max_steps=12)
        first, last = sharding.shard_bounds(total, world_size, rank)
        assert env.first_env == first and env.num_envs == last - first
        env.reset(seed=77)
        res = _rollout(env, tape[:, first:last])
        mx = sharding.max_over_ranks([float(rank + 1), 5.0], "cpu")
        sm = sharding.sum_over_ranks([float(env.num_envs)], "cpu")
        q.put((rank, first, last, res, mx, sm))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gloo_ranks_equal_one_rank(monkeypatch):
    total, T = 22, 30
    tape = np.random.default_rng(3).integers(0, 7, (T, total, 4)).astype(np.int8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, tape, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference run of the same global batch
    import multigrid_b200.env as env_mod
    from multigrid_b200.envs import make
    from tests.hostsim.fake_engine import HostSimStepEngine
    monkeypatch.setattr(env_mod, "StepEngine", HostSimStepEngine)
    env = make("MultiGrid-Empty-8x8-v0", agents=4, num_envs=total, device="cpu", auto_reset=True, max_steps=12)
    env.reset(seed=77)
    ref = _rollout(env, tape)
    for rank, first, last, res, mx, sm in got:
        assert mx == [2.0, 5.0] and sm == [float(total)]
        for t in range(T):
            for a, b in zip(res[t], ref[t]):
                np.testing.assert_array_equal(a, b[first:last], err_msg=f"rank {rank} step {t}")
