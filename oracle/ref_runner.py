#!/usr/bin/env python
"""Time the UNMODIFIED reference's own CPU path (MultiGridEnv.step, multigrid/base.py:303-346: Python +
numba) on this box's host cores. TEST / MEASUREMENT INFRASTRUCTURE (bench.py's `--impl reference` arm and
`cpu_baseline` leg run it as a subprocess; nothing in the product imports it).

    python oracle/ref_runner.py --env MultiGrid-Empty-8x8-v0 --agents 4 --view 7 --procs 32 --seconds 15

BASELINE.md section 3: P worker processes (fork), each ONE reference env instance built by the reference's own
`gym.make(id, agents=n, agent_view_size=V)`, 200 warm-up steps (numba JIT + caches), then uniform random actions
over the 7 actions with reset when all agents are terminated or the episode is truncated; 1 agent-step = one
agent slot of one env advanced by one step(). Prints one JSON object: the sum over processes, the per-process
rates, the process count and the CPU model. The reference lives under oracle/_ref/multigrid (oracle/make_ref.py);
gymnasium / aenum / pygame are the stand-ins of tests/golden/shims/.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
SHIMS = os.path.join(ROOT, "tests", "golden", "shims")


def available() -> str | None:
    """None when the staged reference can run here, else the reason it cannot."""
    if not os.path.isdir(os.path.join(REF_DIR, "multigrid")):
        return "oracle/_ref/multigrid is not staged (python oracle/make_ref.py in the build container)"
    try:
        import numba  # noqa: F401
    except Exception as exc:  # noqa: BLE001
        return f"numba is not importable: {exc}"
    return None


def _make_env(env_id, agents, view):
    sys.path.insert(0, REF_DIR)
    sys.path.insert(0, SHIMS)
    import gymnasium as gym  # shim
    import multigrid.envs  # noqa: F401  (the reference; registers the ids)
    return gym.make(env_id, agents=agents, agent_view_size=view)


def _episode_loop(env, rng, n, steps=None, seconds=None):
    """Random-action stepping with reset on all-terminated / truncated; returns (env steps, seconds)."""
    import numpy as np  # noqa: F401
    done_steps, t0 = 0, time.perf_counter()
    while True:
        block = 64 if steps is None else min(64, steps - done_steps)
        acts = rng.integers(0, 7, size=(block, n))
        for t in range(block):
            obs, rew, term, trunc, _ = env.step(dict(enumerate(acts[t].tolist())))
            if all(term.values()) or all(trunc.values()):
                env.reset()
        done_steps += block
        dt = time.perf_counter() - t0
        if (steps is not None and done_steps >= steps) or (seconds is not None and dt >= seconds):
            return done_steps, dt


def _worker(rank, env_id, agents, view, warm, steps, seconds, cache_dir, barrier, out):
    import numpy as np
    os.environ["NUMBA_CACHE_DIR"] = cache_dir
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    env = _make_env(env_id, agents, view)
    env.reset(seed=1000 + rank)
    rng = np.random.default_rng(rank)
    _episode_loop(env, rng, agents, steps=warm)
    barrier.wait()  # every process is warm (JIT done) before any of them is timed
    done, dt = _episode_loop(env, rng, agents, steps=steps, seconds=seconds)
    out.put((rank, done, dt))


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run(env_id, agents, view, procs, warm=200, steps=None, seconds=10.0, cache_dir=None):
    why = available()
    if why:
        return {"unavailable": why}
    cache_dir = cache_dir or os.path.join("/tmp", f"numba_cache_ref_{os.getuid()}")
    os.makedirs(cache_dir, exist_ok=True)
    ctx = mp.get_context("fork")
    # one process first: it fills the on-disk numba cache, so the others load instead of compiling P times
    t_jit = time.perf_counter()
    out, bar = ctx.Queue(), ctx.Barrier(1)
    p0 = ctx.Process(target=_worker, args=(0, env_id, agents, view, warm, 64, None, cache_dir, bar, out))
    p0.start(); p0.join()
    if p0.exitcode != 0:
        return {"unavailable": f"the reference failed to run here (exit code {p0.exitcode})"}
    out.get()
    jit_s = time.perf_counter() - t_jit
    out, bar = ctx.Queue(), ctx.Barrier(procs)
    ps = [ctx.Process(target=_worker, args=(r, env_id, agents, view, warm, steps, seconds, cache_dir, bar, out))
          for r in range(procs)]
    for p in ps:
        p.start()
    res = [out.get() for _ in ps]
    for p in ps:
        p.join()
    rates = [agents * d / dt for _, d, dt in sorted(res)]
    return {
        "agent_steps_per_s": sum(rates), "processes": procs, "per_process": [round(r) for r in rates],
        "env_steps_per_process": sorted(d for _, d, _ in res)[len(res) // 2],
        "seconds": max(dt for _, _, dt in res), "jit_warm_seconds": round(jit_s, 1), "cpu_model": cpu_model(),
        "env": env_id, "agents": agents, "view": view,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="MultiGrid-Empty-8x8-v0")
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--view", type=int, default=7)
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--warm", type=int, default=200)
    ap.add_argument("--steps", type=int, default=None, help="env steps per process (default: run for --seconds)")
    ap.add_argument("--seconds", type=float, default=10.0)
    args = ap.parse_args()
    print(json.dumps(run(args.env, args.agents, args.view, args.procs, args.warm, args.steps,
                         None if args.steps else args.seconds)), flush=True)


if __name__ == "__main__":
    main()
