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
