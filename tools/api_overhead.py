#!/usr/bin/env python
"""Host-side cost of the public API per step (device-resident actions, no host sync inside the loop):
StepEngine.step (ctypes call) and BatchedMultiGridEnv.step (per-agent dicts of views) against the kernel time.
    python tools/api_overhead.py [--envs 65536]"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_b200.envs import make  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--steps", type=int, default=3000)
ap.add_argument("--env-id", default="MultiGrid-Empty-8x8-v0")
ap.add_argument("--agents", type=int, default=4)
args = ap.parse_args()
env = make(args.env_id, agents=args.agents, num_envs=args.envs, device="cuda:0", auto_reset=True, layout_seed=0)
env.reset(seed=0)
acts = torch.randint(0, 7, (64, args.envs, args.agents), device="cuda:0", dtype=torch.int32).to(torch.int8)
for name, fn in (("engine.step", lambda a: env.engine.step(a)), ("env.step(tensor)", lambda a: env.step(a)),
                 ("env.step(tensor, chained=True)", lambda a: env.step(a, chained=True)),
                 ("env.step(dict)", lambda a: env.step({i: a[:, i] for i in range(args.agents)}))):
    for k in range(50):
        fn(acts[k % 64])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        fn(acts[k % 64])
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(json.dumps(dict(env=args.env_id, call=name, envs=args.envs, host_us_per_call=round(1e6 * t_issue / args.steps, 2),
                          wall_us_per_step=round(1e6 * t_all / args.steps, 2),
                          gagent_steps_s=round(args.envs * args.agents * args.steps / t_all / 1e9, 3))), flush=True)
