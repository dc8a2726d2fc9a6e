#!/usr/bin/env python
"""Per-warp phase timeline of ONE steady-state launch of the fused kernel (diagnostics).

    python tools/trace_timeline.py [--config empty8]

Uses mg_debug_set_trace: each warp records %globaltimer at start / after load / after step /
after obs / end. Prints the distribution of phase durations and of start/end times relative to
the first warp, i.e. how much of the launch is exposed load latency, tail, etc.
"""
import argparse, json, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
TRACE_LIB = os.path.join(ROOT, "multigrid_b200", "_lib", "libmultigrid_b200_trace.so")
_SRC = [os.path.join(ROOT, "multigrid_b200", "csrc", f) for f in ("mg_cabi.cu", "mg_kernels.cuh", "mg_static.cuh")] + \
       [os.path.join(ROOT, "include", "multigrid_b200.h")]
if not os.path.exists(TRACE_LIB) or any(os.path.getmtime(f) > os.path.getmtime(TRACE_LIB) for f in _SRC):
    subprocess.check_call([sys.executable, "-m", "multigrid_b200.build", "--trace"], cwd=ROOT)
os.environ["MG_LIB"] = TRACE_LIB
import bench, kbench  # noqa: E402
from multigrid_b200 import _cabi  # noqa: E402
from multigrid_b200.engine import EngineConfig, StepEngine  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--config", default="empty8"); ap.add_argument("--group", type=int, default=16)
ap.add_argument("--rollout", type=int, default=0, help="trace one mg_rollout launch of this many steps instead")
ap.add_argument("--chained", action="store_true", help="trace one launch in the middle of a run of chained launches")
args = ap.parse_args()
W, H, n, V, E, max_steps, _ = kbench.CONFIGS[args.config]
dev = torch.device("cuda", 0); lib = _cabi.load()
cfg = EngineConfig(width=W, height=H, num_agents=n, view_size=V, max_steps=max_steps, auto_reset=True, stream_state=True)
pg, pa = kbench.layout(W, H, n)
engines = []
for r in range(8):
    eng = StepEngine(cfg, E, dev, pg, pa)
    st, inc = bench.pcg_words(r * E, E)
    eng.load_state(pcg_state=st, pcg_inc=inc)
    eng.reset_from_pool()
    engines.append(eng)
tape = torch.randint(0, 7, (32, E, n), device=dev, dtype=torch.int32).to(torch.int8)
for k in range(96):
    engines[k % 8].step(tape[k % 32])
torch.cuda.synchronize()
groups = (E + 7) // 8  # (enough rows for any group size; unused rows stay zero and are dropped below)
buf = torch.zeros((groups, 8), dtype=torch.int64, device=dev)
if args.rollout:
    T = args.rollout
    rt = torch.randint(0, 7, (T, E, n), device=dev, dtype=torch.int32).to(torch.int8)
    out = engines[0].rollout(rt)
    torch.cuda.synchronize()
    lib.mg_debug_set_trace(buf.data_ptr())
    engines[0].rollout(rt, out)
    torch.cuda.synchronize()
    lib.mg_debug_set_trace(None)
    t = buf.cpu().numpy().astype(np.float64)
    t = t[t[:, 0] > 0]
    print(f"rollout T={T}: launch span {(t[:, 4].max() - t[:, 0].min()) / 1e3:.2f} us = {(t[:, 4].max() - t[:, 0].min()) / 1e3 / T:.2f} us/step")
    print("last iteration of every warp (slots: 5 iteration start, 1 loaded, 2 stepped, 3 observed; 6 = fence after the previous iteration):")
    for a, b, nm in [(5, 1, "load wait"), (1, 2, "reset+step"), (2, 3, "obs"), (6, 5, "loop edge"), (5, 3, "iteration"), (3, 4, "store+drain")]:
        d = (t[:, b] - t[:, a]) / 1e3
        q = np.percentile(d, [0, 10, 50, 90, 100])
        print(f"{nm:11s} dur(us) min/p10/p50/p90/max = " + " ".join(f"{v:7.2f}" for v in q) + f"  mean {d.mean():.2f}")
    life = (t[:, 4] - t[:, 0]) / 1e3
    print("warp life us min/p50/max:", life.min(), np.median(life), life.max())
    sys.exit(0)
if args.chained:
    for k in range(12):  # launches 0..11 on rotating engines, all chained; launch 6 is traced
        if k == 6:
            lib.mg_debug_set_trace(buf.data_ptr())
        engines[k % 8].step(tape[k % 32], chained=True)
        if k == 6:
            lib.mg_debug_set_trace(None)
    torch.cuda.synchronize()
else:
    lib.mg_debug_set_trace(buf.data_ptr())
    engines[0].step(tape[0])
    torch.cuda.synchronize()
    lib.mg_debug_set_trace(None)
t = buf.cpu().numpy().astype(np.float64)
t = t[t[:, 0] > 0]  # rows of warps that ran (the buffer is sized for the smallest group size)
groups = len(t)
t0 = t[:, 0].min()
rel = (t[:, :5] - t0) / 1e3  # us
names = ["start", "loaded", "stepped", "observed", "end"]
print(f"launch span: {rel[:, 4].max():.2f} us, groups {groups}")
for i, nm in enumerate(names):
    q = np.percentile(rel[:, i], [0, 10, 50, 90, 100])
    print(f"{nm:9s} t(us) min/p10/p50/p90/max = " + " ".join(f"{v:7.2f}" for v in q))
for a, b, nm in [(0, 1, "load wait"), (1, 2, "reset+step"), (2, 3, "obs"), (3, 4, "store+drain"), (0, 4, "warp life")]:
    d = rel[:, b] - rel[:, a]
    q = np.percentile(d, [0, 10, 50, 90, 100])
    print(f"{nm:11s} dur(us) min/p10/p50/p90/max = " + " ".join(f"{v:7.2f}" for v in q) + f"  mean {d.mean():.2f}")
# concurrency over time
edges = np.linspace(0, rel[:, 4].max(), 25)
act = [(np.sum((rel[:, 0] <= x) & (rel[:, 4] > x)), np.sum((rel[:, 0] <= x) & (rel[:, 1] > x))) for x in edges]
print("t(us): resident warps / of which waiting for load")
print(" ".join(f"{x:.1f}:{a}/{b}" for x, (a, b) in zip(edges, act)))
sm = t[:, 7].astype(int)
print("warps per SM min/max:", np.bincount(sm, minlength=148).min(), np.bincount(sm, minlength=148).max())
