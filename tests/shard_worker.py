"""Worker of tests/test_sharding_gpu.py, launched by torch.distributed.run: this rank's shard of one global env batch
on the real CUDA engine, stepped T times on a shared action tape; every rank returns a 64-bit checksum per GLOBAL
env id of everything a step returns (obs, reward, terminated, truncated) plus the final state. With fewer GPUs than
ranks the ranks share cuda:0 (the equivalence is about global env ids, not about devices)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multigrid_b200 import sharding  # noqa: E402


def checksums(env, tape, first, last):
    E = last - first
    dev = env.device
    h = torch.zeros(E, dtype=torch.int64, device=dev)
    eng = env.engine
    wts = None
    for t in range(tape.shape[0]):
        a = torch.from_numpy(tape[t, first:last]).to(dev)
        env.step(a)
        obs = eng.obs.reshape(E, -1).to(torch.int64)
        if wts is None:
            wts = torch.arange(1, obs.shape[1] + 1, dtype=torch.int64, device=dev) * 2654435761
        x = (obs * wts).sum(1)
        x = x * 31 + eng.reward.view(torch.int64).sum(1)
        x = x * 31 + (eng.terminated.to(torch.int64) << torch.arange(eng.terminated.shape[1], device=dev)).sum(1)
        x = x * 31 + eng.truncated.to(torch.int64)
        h = h * 1000003 + x
    h = h * 1000003 + eng.agents.view(E, -1).to(torch.int64).mul(wts[: eng.agents[0].numel()]).sum(1)
    h = h * 1000003 + eng.step_count.to(torch.int64) * 7 + eng.pcg_state.sum(1)
    env.check()
    return h.cpu().numpy()


def main():
    env_id, agents, total, T, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    auto_reset = sys.argv[6] == "1"  # (a cycling layout POOL is per shard; fresh per-env layouts are not: see DESIGN)
    rank, world, local_rank = sharding.world()
    ngpu = torch.cuda.device_count()
    dev = f"cuda:{local_rank if local_rank < ngpu else 0}"
    if world > 1:
        dist.init_process_group("gloo")
    tape = np.random.default_rng(11).integers(0, 7, (T, total, agents)).astype(np.int8)
    first, last = sharding.shard_bounds(total, world, rank)
    env = sharding.make_sharded(env_id, total, agents=agents, device=dev, auto_reset=auto_reset, max_steps=40, layout_seed=5,
                                pool_size=total)
    assert env.first_env == first and env.num_envs == last - first
    env.reset(seed=123)
    h = checksums(env, tape, first, last)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (first, last, h))
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered = [(first, last, h)]
    if rank == 0:
        full = np.zeros(total, np.int64)
        for f, l, hh in gathered:
            full[f:l] = hh
        np.save(out, full)


if __name__ == "__main__":
    main()
