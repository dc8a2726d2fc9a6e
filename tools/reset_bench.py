#!/usr/bin/env python
"""reset() wall time of a random-layout env with the layout pool generated on the device
(mg_gen_layouts_empty_random) vs in Python (multigrid_b200/layouts.py).
    python tools/reset_bench.py [--envs 65536] [--pool 4096]"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_b200.envs import make  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--pool", type=int, default=4096)
args = ap.parse_args()
ap2 = [("MultiGrid-Empty-Random-6x6-v0", 4), ("MultiGrid-BlockedUnlockPickup-v0", 2)]
for env_id, n, dl in [(i, n, d) for i, n in ap2 for d in (True, False)]:
    env = make(env_id, agents=n, num_envs=args.envs, device="cuda:0", layout_seed=1,
               pool_size=args.pool, device_layouts=dl)
    env.reset(seed=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.reset(seed=1)
    torch.cuda.synchronize()
    print(json.dumps(dict(env=env_id, device_layouts=dl, envs=args.envs, pool=args.pool,
                          reset_s=round(time.perf_counter() - t0, 4))), flush=True)

# fresh layouts on auto-reset (mg_refresh_done_layouts after every step) vs the cycling pool: us per step
for env_id, n, E in [("MultiGrid-Empty-Random-6x6-v0", 4, 65536), ("MultiGrid-BlockedUnlockPickup-v0", 2, 32768)]:
    for fresh in (False, True):
        env = make(env_id, agents=n, num_envs=E, device="cuda:0", layout_seed=1, auto_reset=True, fresh_layouts=fresh,
                   max_steps=64)
        env.reset(seed=0)
        g = torch.Generator(device="cuda:0").manual_seed(0)
        tape = torch.randint(0, 7, (64, E, n), device="cuda:0", dtype=torch.int32, generator=g).to(torch.int8)
        for k in range(130):
            env.step(tape[k % 64])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 640  # 10 episodes of 64 steps: every env resets (and, with fresh=True, gets a new layout) 10 times
        for k in range(K):
            env.step(tape[k % 64])
        torch.cuda.synchronize()
        env.check()
        print(json.dumps(dict(env=env_id, envs=E, fresh_layouts=fresh, us_per_step=round(1e6 * (time.perf_counter() - t0) / K, 2))),
              flush=True)
