#!/usr/bin/env python
"""Kernel-only timing sweep of the fused step+observe kernel over the launch knobs
(MG_GROUP, MG_WPB, MG_NO_BULK, ...) and over the BASELINE.json configurations.

    python tools/kbench.py [--configs empty8,bup,empty16] [--steps 200]

Not the bench contract (that is bench.py): a development tool to pick launch geometry.
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multigrid_b200.engine import EngineConfig, StepEngine  # noqa: E402

CONFIGS = {
    # name: (W, H, n, V, E, max_steps, mutable_grid)
    "empty8": (8, 8, 4, 7, 65536, 256, False),
    "bup": (11, 6, 2, 7, 32768, 576, True),
    "empty16": (16, 16, 8, 9, 16384, 1024, False),
    "empty8h": (8, 8, 4, 7, 32768, 256, False),   # half batch: launch-plan heuristics
    "empty8q": (8, 8, 4, 7, 8192, 256, False),
}


def layout(W, H, n):
    grid = np.zeros((1, W, H, 3), np.int8)
    grid[..., 0] = 1
    grid[0, 0, :] = grid[0, W - 1, :] = (2, 5, 0)
    grid[0, :, 0] = grid[0, :, H - 1] = (2, 5, 0)
    grid[0, W - 2, H - 2] = (8, 1, 0)
    agents = np.zeros((1, n, 8), np.int8)
    agents[..., 1] = 1
    agents[..., 2] = 1
    agents[..., 4] = 1
    agents[..., 7] = np.arange(n) % 6
    return grid, agents


def time_config(name, steps, knobs, replicas=8):
    W, H, n, V, E, max_steps, mutable = CONFIGS[name]
    dev = torch.device("cuda", 0)
    cfg = EngineConfig(width=W, height=H, num_agents=n, view_size=V, max_steps=max_steps, auto_reset=True,
                       stream_state=bool(int(os.environ.get("KB_STREAM", "1"))))
    pg, pa = layout(W, H, n)
    engines = []
    for r in range(replicas):
        eng = StepEngine(cfg, E, dev, pg, pa)
        st, inc = bench.pcg_words(r * E, E)
        eng.load_state(pcg_state=st, pcg_inc=inc)
        eng.reset_from_pool()
        engines.append(eng)
    gen = torch.Generator(device=dev).manual_seed(7)
    NT = 256  # long tape: a short action cycle keeps agents near their start cells and flatters the kernel
    tape = torch.randint(0, 7, (NT, E, n), generator=gen, device=dev, dtype=torch.int32).to(torch.int8)
    for k in range(NT):
        engines[k % replicas].step(tape[k % NT])
    torch.cuda.synchronize()
    bpe = bench.algorithmic_bytes_per_env_step(W, H, n, V, mutable)
    out = []
    # every knob set is timed from the SAME state (all envs of a fixed-start layout share their episode
    # phase: how far the agents have spread, and with it the cost of a launch, drifts with the step count)
    names = ("cells", "agents", "step_count", "pcg_state", "layout_idx", "hook_state", "chain")
    snap = [{k: getattr(e, k).clone() for k in names} for e in engines]
    for knob in knobs:
        for key in ("MG_GROUP", "MG_WPB", "MG_NO_BULK", "MG_GENERIC_VIEW", "MG_PDL", "MG_L2HINT", "MG_NO_DEDUP", "CHAINED",
                    "NOSTATIC", "REPLICAS", "ONEHOT"):
            os.environ.pop(key, None)
        os.environ.update({k: str(v) for k, v in knob.items()})
        chained = bool(int(os.environ.get("CHAINED", "0")))  # (a kbench knob, not a library one)
        for e in engines:  # NOSTATIC=1 (a kbench knob): the general kernel on a static-grid batch
            e.use_static = not int(os.environ.get("NOSTATIC", "0"))
        nrep = int(os.environ.get("REPLICAS", str(replicas)))  # REPLICAS=1: one batch stepped closed-loop (L2-resident)
        # ONEHOT (a kbench knob): 1 = the step kernel also writes the one-hot images (MgStepOut.one_hot),
        # 2 = a separate mg_one_hot pass after every step (what the fused image replaces)
        onehot = int(os.environ.get("ONEHOT", "0"))
        for e in engines:
            if onehot:
                e.enable_one_hot()
            e._oh_keep = e.one_hot if e.one_hot is not None else getattr(e, "_oh_keep", None)
            e.one_hot = e._oh_keep if onehot == 1 else None
            e._c = None      # rebuild the structs (one_hot pointer) and
            e._plans = {}    # the prepared launches (they captured the previous knobs)
        for e, sn in zip(engines, snap):
            for k in names:
                getattr(e, k).copy_(sn[k])
        torch.cuda.synchronize()
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            for k in range(4):
                engines[k % replicas].step(tape[k % NT])
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for k in range(steps):
                    e = engines[k % nrep]
                    e.step(tape[k % NT], chained=chained)
                    if onehot == 2:
                        e.lib.mg_one_hot(V, E * n, e.obs_stride, e.obs_buf.data_ptr(), e._oh_keep.data_ptr(),
                                         stream.cuda_stream)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                ev0.record(stream)
                graph.replay()
                ev1.record(stream)
            torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
        us = 1e3 * best / steps
        rec = dict(config=name, **knob, us_per_launch=round(us, 2),
                   gagent_steps_s=round(E * n / us / 1e3, 2), gbs=round(bpe * E / us / 1e3, 1))
        print(json.dumps(rec), flush=True)
        out.append(rec)
    return out


def time_rollout(name, T, knobs):
    """mg_rollout: T steps in one launch on one engine (state stays L2/on-chip resident)."""
    W, H, n, V, E, max_steps, mutable = CONFIGS[name]
    dev = torch.device("cuda", 0)
    cfg = EngineConfig(width=W, height=H, num_agents=n, view_size=V, max_steps=max_steps, auto_reset=True,
                       stream_state=True)
    pg, pa = layout(W, H, n)
    eng = StepEngine(cfg, E, dev, pg, pa)
    st, inc = bench.pcg_words(0, E)
    eng.load_state(np.repeat(pg, E, 0), np.repeat(pa, E, 0), None, st, inc, None)
    gen = torch.Generator(device=dev).manual_seed(7)
    tape = torch.randint(0, 7, (T, E, n), generator=gen, device=dev, dtype=torch.int32).to(torch.int8)
    out = eng.rollout(tape)  # warm-up: spreads the agents, allocates the outputs
    for _ in range(max(1, 256 // T)):
        eng.rollout(tape, out)
    torch.cuda.synchronize()
    bpe = bench.rollout_bytes_per_env_step(n, V)
    for knob in knobs:
        for key in ("MG_GROUP", "MG_WPB", "MG_NO_BULK", "MG_GENERIC_VIEW", "MG_PDL", "MG_L2HINT"):
            os.environ.pop(key, None)
        os.environ.update({k: str(v) for k, v in knob.items()})
        eng.rollout(tape, out)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            eng.rollout(tape, out)
            ev1.record()
            torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
        us = 1e3 * best / T
        print(json.dumps(dict(config=name, rollout_T=T, **knob, us_per_step=round(us, 2),
                              gagent_steps_s=round(E * n / us / 1e3, 2), gbs=round(bpe * E / us / 1e3, 1))), flush=True)


def time_one_hot(V=7, A=262144):
    """mg_one_hot on the bench batch's observations: output 21 B per cell, HBM-write-bound at best."""
    import ctypes as C
    from multigrid_b200 import _cabi
    lib = _cabi.load()
    stride = _cabi.obs_agent_stride(V)
    obs = torch.randint(0, 4, (A, stride), device="cuda", dtype=torch.int32).to(torch.int8)
    outs = [torch.empty((A, V, V, 21), dtype=torch.uint8, device="cuda") for _ in range(3)]  # 3 x 270 MB > L2
    for name, env in (("tile", "0"), ("v16", "v16"), ("w32", "1")):
        os.environ["MG_ONE_HOT_W32"] = "1" if env == "1" else "0"
        os.environ["MG_ONE_HOT_V16"] = "1" if env == "v16" else "0"
        for o in outs:
            lib.mg_one_hot(V, A, stride, obs.data_ptr(), o.data_ptr(), None)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for k in range(30):
            lib.mg_one_hot(V, A, stride, obs.data_ptr(), outs[k % 3].data_ptr(), None)
        ev1.record()
        torch.cuda.synchronize()
        us = 1e3 * ev0.elapsed_time(ev1) / 30
        nbytes = A * V * V * 21 + A * stride
        print(json.dumps(dict(kernel=f"one_hot_{name}", agents=A, V=V, us=round(us, 2),
                              gbs=round(nbytes / us / 1e3, 1))), flush=True)
    os.environ.pop("MG_ONE_HOT_W32", None)
    os.environ.pop("MG_ONE_HOT_V16", None)
    # mg_obs_features: the float32 23-channel network input straight from the observations
    dirs = torch.randint(0, 4, (A, 8), device="cuda", dtype=torch.int32).to(torch.int8)
    lut = torch.stack([torch.cos(2 * torch.pi * torch.arange(4) / 4), torch.sin(2 * torch.pi * torch.arange(4) / 4)], -1).cuda()
    fouts = [torch.empty((A, V, V, 23), dtype=torch.float32, device="cuda") for _ in range(2)]
    for o in fouts:
        lib.mg_obs_features(V, A, stride, obs.data_ptr(), dirs.data_ptr(), 8, lut.data_ptr(), o.data_ptr(), None)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(20):
        lib.mg_obs_features(V, A, stride, obs.data_ptr(), dirs.data_ptr(), 8, lut.data_ptr(), fouts[k % 2].data_ptr(), None)
    ev1.record()
    torch.cuda.synchronize()
    us = 1e3 * ev0.elapsed_time(ev1) / 20
    nbytes = A * V * V * 23 * 4 + A * stride
    print(json.dumps(dict(kernel="obs_features", agents=A, V=V, us=round(us, 2), gbs=round(nbytes / us / 1e3, 1))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="empty8")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--groups", default="0")
    ap.add_argument("--wpbs", default="0")
    ap.add_argument("--nobulk", default="0")
    ap.add_argument("--extra", default="", help="comma list of extra KEY=VAL knob sets to also try, e.g. MG_PDL=1,MG_PDL=1+MG_L2HINT=3")
    ap.add_argument("--one-hot", action="store_true", help="time the one-hot kernels on the bench batch")
    ap.add_argument("--rollout", type=int, default=0, help="also time mg_rollout with this many steps per launch")
    args = ap.parse_args()
    knobs = []
    for g, w, nb in itertools.product(args.groups.split(","), args.wpbs.split(","), args.nobulk.split(",")):
        k = {}
        if int(g):
            k["MG_GROUP"] = int(g)
        if int(w):
            k["MG_WPB"] = int(w)
        if int(nb):
            k["MG_NO_BULK"] = 1
        knobs.append(k)
        for kv in [x for x in args.extra.split(",") if x]:
            knobs.append({**k, **dict(item.split("=") for item in kv.split("+"))})
    if args.one_hot:
        time_one_hot()
    for name in args.configs.split(","):
        if args.steps > 0:
            time_config(name, args.steps, knobs)
        if args.rollout:
            time_rollout(name, args.rollout, knobs)


if __name__ == "__main__":
    main()
