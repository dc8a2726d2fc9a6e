#!/usr/bin/env python
"""Benchmark of the MultiGrid step/observe hot path on B200 (contract: see DESIGN.md, Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config empty8|bup|empty16]   # this repo's CUDA engine
    python bench.py --impl reference [--steps K] [--warmup W] [--config ...]             # the reference on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...                                # one rank per GPU, env axis sharded

Workloads (BASELINE.json `configs`, per GPU; default = the configuration the metric is quoted on):
    empty8   configs[1]  MultiGrid-Empty-8x8-v0, agents=4, view 7, num_envs=65536
    bup      configs[2]  MultiGrid-BlockedUnlockPickup-v0, agents=2, view 7, num_envs=32768
    empty16  configs[3]  MultiGrid-Empty-16x16-v0, agents=8, view 9, num_envs=16384
uniform random actions over the 7 actions, "next-step" auto-reset; a step is one fused mg_step_obs launch over the
whole batch. Metric: agent-steps/s (1 agent-step = one agent slot of one env advanced by one step; terminated /
skipped agents count, on CPU and GPU alike).

`value` is measured with PLAIN launches (each launch waits for the whole previous one: what a closed-loop caller
gets), K of them captured in one CUDA graph over rotating state replicas so that every launch finds its inputs
in HBM, not L2. One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent-steps/sec on Empty-8x8 agents=4 num_envs=65536; HBM GB/s vs 8 TB/s peak"
CONFIGS = {
    # W, H, max_steps as the reference env classes set them (envs/empty.py:145, envs/blockedunlockpickup.py:132)
    "empty8": dict(env_id="MultiGrid-Empty-8x8-v0", agents=4, view=7, envs=65536, W=8, H=8, max_steps=256,
                   mutable_grid=False, baseline="BASELINE.json configs[1]; configs[4] at 8 GPUs"),
    "bup": dict(env_id="MultiGrid-BlockedUnlockPickup-v0", agents=2, view=7, envs=32768, W=11, H=6, max_steps=576,
                mutable_grid=True, baseline="BASELINE.json configs[2]"),
    "empty16": dict(env_id="MultiGrid-Empty-16x16-v0", agents=8, view=9, envs=16384, W=16, H=16, max_steps=1024,
                    mutable_grid=False, baseline="BASELINE.json configs[3]"),
}
REPLICAS = 8   # state replicas rotated through so each launch finds its inputs in HBM, not L2
BURN_IN = 64   # untimed steps that de-synchronise the envs before anything is measured
VERIFY_ENVS = 2048  # envs of a replica replayed on the C oracle after the timed region
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used if MEASURED_PEAKS.json is absent
REF_CHUNK = 4096  # --impl reference: one bench "step" = this many env steps on each worker process


def algorithmic_bytes_per_env_step(W, H, n, V, mutable_grid=False):
    """packed state round trip + outputs, per env-step."""
    reads = 3 * W * H + 7 * n + 2 + 16 + 16 + n
    writes = (3 * W * H if mutable_grid else 0) + 7 * n + 2 + 16 + 3 * n * V * V + 8 * n + n + 1
    return reads + writes


def rollout_bytes_per_env_step(n, V):
    """SURVEY.md section 8(d), in-kernel multi-step rollout: state stays on chip, so the per-env-step floor
    is actions in + outputs out (obs, f64 reward, terminated, truncated, direction)."""
    return n + 3 * n * V * V + 8 * n + n + 1 + n


def moved_bytes_per_env_step(W, H, n, ostride, static):
    """Bytes the launch really moves through HBM per env-step with the engine's layout:
    8-byte agent records, int32 counters, obs slots of `ostride` bytes; the general kernel also reads the env's
    padded 4-byte cell words and its layout cursor, the static-grid kernel reads no grid at all."""
    reads = 8 * n + 4 + 16 + 16 + n
    writes = 8 * n + 4 + 16 + n * ostride + 8 * n + n + 1
    if not static:
        reads += 4 * (W + 1) * (H + 1) + 4 + 16  # cells, layout cursor, dirty flag record
        writes += 4
    return reads + writes


def empty_layout(size, n):
    """EmptyEnv._gen_grid, fixed start (envs/empty.py:151-170), packed engine layout."""
    grid = np.zeros((1, size, size, 3), np.int8)
    grid[..., 0] = 1
    for sl in (np.s_[0, 0, :], np.s_[0, size - 1, :], np.s_[0, :, 0], np.s_[0, :, size - 1]):
        grid[sl] = (2, 5, 0)
    grid[0, size - 2, size - 2] = (8, 1, 0)
    agents = np.zeros((1, n, 8), np.int8)
    agents[..., 1] = 1
    agents[..., 2] = 1
    agents[..., 4] = 1
    agents[..., 7] = np.arange(n) % 6
    return grid, agents


def pcg_words(first_env, count, base_seed=2024):
    """Per-env numpy PCG64 (state, inc) for global env ids; splitmix-style so it is O(count)."""
    ids = np.arange(first_env, first_env + count, dtype=np.uint64) + np.uint64(base_seed) * np.uint64(1 << 32)

    def mix(x, c):
        x = (x + np.uint64(c)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))

    with np.errstate(over="ignore"):
        st = np.stack([mix(ids, 1), mix(ids, 2)], 1)
        inc = np.stack([mix(ids, 3) | np.uint64(1), mix(ids, 4)], 1)  # PCG increments are odd
    return st, inc


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # NVML missing: clocks are reported as unavailable
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            peaks = json.load(f)
        for key in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if key in peaks:
                return float(peaks[key]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(config):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    capture of this bench command (profiles/r02_traffic.json; written by tools/ncu_traffic.py). None when no
    capture of this configuration is committed: the bench itself cannot read DRAM counters."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            rec = json.load(f).get(config)
        return rec
    except Exception:
        return None


# -------------------------------------------------------------------------------------------------
# CPU side. The reference's own path (oracle/_ref + oracle/ref_runner.py, kind "reference") when it is staged and
# numba imports; the C port of the oracle (oracle/mg_oracle.c, kind "port") otherwise and as a second figure.
# -------------------------------------------------------------------------------------------------
def time_reference_processes(cfg, procs, steps=None, seconds=None, warm=200):
    """The unmodified reference on `procs` worker processes, in a clean subprocess."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "--env", cfg["env_id"],
           "--agents", str(cfg["agents"]), "--view", str(cfg["view"]), "--procs", str(procs), "--warm", str(warm)]
    cmd += ["--steps", str(steps)] if steps else ["--seconds", str(seconds)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"oracle/ref_runner.py failed: {exc}"}


def make_cpu_oracle(cfg, num_envs, nthreads):
    """C oracle port on Empty layouts (the port leg is only quoted for the Empty configurations)."""
    from oracle.c_oracle import COracle, build
    from oracle.mg_oracle import OracleConfig
    build()
    n, size = cfg["agents"], cfg["W"]
    ocfg = OracleConfig(W=size, H=size, n=n, V=cfg["view"], max_steps=cfg["max_steps"], auto_reset=True)
    pg, pa = empty_layout(size, n)
    st, inc = pcg_words(0, num_envs)
    return COracle(ocfg, np.repeat(pg, num_envs, 0), np.repeat(pa, num_envs, 0), st, inc,
                   pool_grid=pg, pool_agents=pa, nthreads=nthreads)


def time_cpu_port(cfg, num_envs, steps, warmup, nthreads, budget_s=None):
    """Returns (agent_steps_per_s, steps_done, seconds)."""
    ora = make_cpu_oracle(cfg, num_envs, nthreads)
    rng = np.random.default_rng(0)
    tape = rng.integers(0, 7, size=(64, num_envs, cfg["agents"])).astype(np.int8)
    for t in range(warmup):
        ora.step(tape[t % 64])
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        ora.step(tape[done % 64])
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return num_envs * cfg["agents"] * done / dt, done, dt


def cpu_baseline(cfg, seconds):
    """cpu_baseline object of the bench line: the reference itself when it can run here, else the port."""
    cores = len(os.sched_getaffinity(0))
    ref = time_reference_processes(cfg, cores, seconds=seconds)
    port = None
    if not cfg["mutable_grid"]:
        v, done, dt = time_cpu_port(cfg, cfg["envs"], 10**9, 2, cores, budget_s=min(seconds, 6.0))
        port = {"value": v, "unit": "agent-steps/s", "cores": cores,
                "sample": f"{done} steps x {cfg['envs']} envs ({dt:.1f} s), oracle/mg_oracle.c with OpenMP"}
    if "unavailable" not in ref:
        out = {"value": ref["agent_steps_per_s"], "unit": "agent-steps/s", "cores": ref["processes"], "kind": "reference",
               "sample": f"the unmodified reference (multigrid/base.py:303-346, Python + numba) in {ref['processes']} "
                         f"processes, one env each, {ref['env_steps_per_process']} env steps per process "
                         f"({ref['seconds']:.1f} s) after 200 warm-up steps; {ref['cpu_model']}",
               "per_process_median": float(np.median(ref["per_process"]))}
        if port:
            out["port"] = port
        return out
    if port is None:
        return {"value": None, "unit": "agent-steps/s", "cores": cores, "kind": "reference", "unavailable": ref["unavailable"]}
    return {**port, "kind": "port", "reference_unavailable": ref["unavailable"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once per box, on rank 0
    cfg = CONFIGS[args.config]
    cores = len(os.sched_getaffinity(0))
    K, W = args.steps, args.warmup
    ref = time_reference_processes(cfg, cores, steps=K * REF_CHUNK, warm=200 + W * REF_CHUNK)
    if "unavailable" not in ref:
        value, kind = ref["agent_steps_per_s"], "reference"
        dt = ref["seconds"]
        sample = (f"the unmodified reference (Python + numba) in {cores} processes, one env each; one bench step = "
                  f"{REF_CHUNK} env steps per process; {K} steps = {ref['env_steps_per_process']} env steps per process in "
                  f"{dt:.1f} s, after 200 + {W} x {REF_CHUNK} warm-up steps; {ref['cpu_model']}")
        note = "the reference's own CPU path, staged under oracle/_ref by oracle/make_ref.py, run by oracle/ref_runner.py"
    else:
        value, done, dt = time_cpu_port(cfg, cfg["envs"], K, W, cores)
        kind, sample = "port", f"{done} steps x {cfg['envs']} envs ({dt:.2f} s), oracle/mg_oracle.c with OpenMP"
        note = f"C port of the oracle: the reference itself cannot run here ({ref['unavailable']})"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.config, args.gpus),
        "note": note,
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, n_gpus):
    c = CONFIGS[name]
    return {
        "workload": f"{c['env_id']} agents={c['agents']} view={c['view']} num_envs={c['envs']} per GPU, uniform random "
                    f"actions, next-step auto-reset ({c['baseline']})",
        "name": name, "num_envs_per_gpu": c["envs"], "num_envs_total": c["envs"] * n_gpus,
        "agents": c["agents"], "view_size": c["view"], "grid": f"{c['W']}x{c['H']}", "max_steps": c["max_steps"],
        "sharding": f"env axis split over {n_gpus} GPU(s), no collective on the step path",
        "l2": f"{REPLICAS} state replicas rotated (each launch touches a batch last used {REPLICAS} launches ago; "
              f"{REPLICAS} x state+outputs > 126 MB L2), no explicit flush",
        "launches": "plain (every launch waits for the whole previous launch), K of them in one CUDA graph",
    }


# -------------------------------------------------------------------------------------------------
# GPU side
# -------------------------------------------------------------------------------------------------
def bind_to_local_cpus(gpu_index):
    """Best effort: run this rank's host-buffer leg on the CPUs next to its GPU (NVML's ideal CPU affinity), so that
    pinned buffers allocated afterwards land on the GPU's NUMA node. A no-op when NVML or the cpuset says no."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001
        pass


def graph_of(stream, torch, fn, K):
    with torch.cuda.stream(stream):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for k in range(K):
                fn(k)
        g.replay()  # untimed: the first replay of an instantiated graph also uploads it to the device
    torch.cuda.synchronize()
    return g


def time_graph(stream, torch, g):
    """Device time of one replay. A ~100 us spin kernel is enqueued first, so that the start event and the graph launch
    are both queued behind it when it ends: the timed window is the K launches on the device, not the host's latency
    between `ev0.record()` and the arrival of the graph launch (under torchrun, N Python processes share the host)."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        torch.cuda._sleep(200_000)
        ev0.record(stream)
        g.replay()
        ev1.record(stream)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1)


def verify_against_oracle(torch, eng, replay, launches_of_replica, tape, n_tape, M):
    """Replays what the timed graph does to one state replica on the C oracle (first M envs) and compares the state
    and the outputs of its last launch: the exact pattern that was timed (rotating replicas, graph, static / dedup /
    streaming cache policy) against the checker. Outside every timed region. Returns a short report."""
    from oracle import mg_oracle as O
    from oracle.c_oracle import COracle
    c = eng.cfg
    ocfg = O.OracleConfig(W=c.width, H=c.height, n=c.num_agents, V=c.view_size, max_steps=c.max_steps,
                          see_through_walls=c.see_through_walls, allow_agent_overlap=c.allow_agent_overlap,
                          joint_reward=c.joint_reward, success_any=c.success_termination_mode == "any",
                          failure_any=c.failure_termination_mode == "any", hook=c.hook, hook_param=c.hook_param,
                          auto_reset=c.auto_reset, layout_stride=c.layout_stride)
    torch.cuda.synchronize()
    W, H = c.width, c.height
    pool_cells = eng.pool_grid.cpu().numpy().view(np.uint32)
    pool_grid = np.stack([pool_cells & 0xff, (pool_cells >> 8) & 0xff, (pool_cells >> 16) & 0xff], -1)[:, :W, :H].astype(np.int8)
    ora = COracle(ocfg, eng.grid[:M].cpu().numpy(), eng.agents[:M].cpu().numpy(),
                  eng.pcg_state[:M].cpu().numpy().view(np.uint64), eng.pcg_inc[:M].cpu().numpy().view(np.uint64),
                  pool_grid=pool_grid, pool_agents=eng.pool_agents.cpu().numpy(),
                  layout_idx=eng.layout_idx[:M].cpu().numpy(), step_count=eng.step_count[:M].cpu().numpy(),
                  nthreads=len(os.sched_getaffinity(0)))
    ora.hook_state[:] = eng.hook_state[:M].cpu().numpy()
    replay()
    torch.cuda.synchronize()
    out = None
    for k in launches_of_replica:
        out = ora.step(tape[k % n_tape][:M].cpu().numpy())
    V = c.view_size
    checks = {
        "grid": np.array_equal(eng.grid[:M].cpu().numpy(), ora.grid),
        "agents": np.array_equal(eng.agents[:M].cpu().numpy(), ora.agents),
        "step_count": np.array_equal(eng.step_count[:M].cpu().numpy(), ora.step_count),
        "pcg_state": np.array_equal(eng.pcg_state[:M].cpu().numpy().view(np.uint64), ora.pcg_state),
        "obs": np.array_equal(eng.obs[:M].cpu().numpy(), out[0]),
        "reward": bool((eng.reward[:M].cpu().numpy() == out[1]).all()),
        "terminated": np.array_equal(eng.terminated[:M].cpu().numpy(), out[2]),
        "truncated": np.array_equal(eng.truncated[:M].cpu().numpy(), out[3]),
    }
    eng.check_status()
    bad = [k for k, ok in checks.items() if not ok]
    if bad:
        raise AssertionError(f"bench verification failed: {bad} differ from the C oracle")
    return {"envs": M, "steps": len(launches_of_replica), "compared": sorted(checks)}


def run_engine(args):
    import torch
    import torch.distributed as dist
    from multigrid_b200 import _cabi
    from multigrid_b200.envs import make

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _cabi.load()
    cfg = CONFIGS[args.config]
    K, Wm = args.steps, max(args.warmup, 3)
    E, n = cfg["envs"], cfg["agents"]

    # REPLICAS independent batches through the public API; seeds are a function of the global env id
    envs = []
    for r in range(REPLICAS):
        env = make(cfg["env_id"], agents=n, agent_view_size=cfg["view"], num_envs=E, device=dev, auto_reset=True,
                   first_env=(rank * REPLICAS + r) * E, layout_seed=7, stream_state=True)
        env.reset(seed=2024)
        envs.append(env)
    engines = [env.engine for env in envs]
    eng = engines[0]
    static = bool(eng._static_ok())

    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    n_tape = max(64, K)  # no action set is replayed inside the timed region (a short cycle keeps agents near
                         # their start cells, which flatters the kernel)
    tape = torch.randint(0, 7, (n_tape, E, n), generator=gen, device=dev, dtype=torch.int32).to(torch.int8)

    def launch(k, chained=False):
        engines[k % REPLICAS].step(tape[k % n_tape], chained=chained)

    for k in range(BURN_IN):
        launch(k)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- `value`: K plain fused launches, inputs resident in HBM, captured in one CUDA graph ----------
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for k in range(Wm):
            launch(k)
        stream.synchronize()
    before = lib.mg_launch_count()
    graph = graph_of(stream, torch, lambda k: launch(Wm + k), K)
    launches = (lib.mg_launch_count() - before)  # (capture issues each launch once)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = time_graph(stream, torch, graph)
    if world > 1:
        dist.barrier()

    # ---- informational: closed loop on ONE batch (state + outputs stay in L2), and chained launches where the
    # general kernel runs (MG_FLAG_CHAINED; the static-grid kernel has no chained form)
    graph_cl = graph_of(stream, torch, lambda k: engines[0].step(tape[(Wm + k) % n_tape]), K)
    ms_closed = time_graph(stream, torch, graph_cl)
    ms_chained = None
    if not static:
        graph_ch = graph_of(stream, torch, lambda k: launch(Wm + k, chained=True), K)
        ms_chained = time_graph(stream, torch, graph_ch)

    # ---- verification of what was timed: two replicas of the timed graph against the C oracle -------
    verified = None
    if not args.no_verify and rank == 0:
        verified = []
        for r in (0, REPLICAS - 1):
            ks = [Wm + k for k in range(K) if (Wm + k) % REPLICAS == r]
            verified.append(dict(replica=r, **verify_against_oracle(
                torch, engines[r], graph.replay, ks, tape, n_tape, min(VERIFY_ENVS, E))))

    # ---- `e2e`: the host-buffer call: H2D actions from pinned memory, kernel, D2H results into pinned memory.
    # Headline = mg_step_obs_host_packed (observations cross PCIe in the 9-bit-per-cell wire format, decoded on the
    # host by engine.unpack_obs); the unpacked call (mg_step_obs_host, raw 3-byte cells) is timed beside it.
    bind_to_local_cpus(local_rank)
    # the caller's actions of each step live in pinned host memory (8 different sets, cycled)
    host_tape = [tape[k].cpu().pin_memory() for k in range(8)]
    K2 = max(3, min(K, 50))

    def time_host(packed):
        h = eng.host_buffers(packed)
        for k in range(3):
            eng.step_host(packed=packed, actions=host_tape[k % 8])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(K2):  # H2D of this step's actions, kernel(s), D2H; results are in pinned host memory on return
            eng.step_host(synchronize=True, packed=packed, actions=host_tape[k % 8])
        return time.perf_counter() - t0, h

    e2e_raw_s, h = time_host(False)
    e2e_p9_s, h = time_host(True)
    from multigrid_b200.engine import unpack_obs
    M = min(VERIFY_ENVS, E)  # the packed result decodes to exactly the device-side observations
    assert np.array_equal(unpack_obs(h["obs_packed"][:M], cfg["view"]), eng.obs[:M].cpu().numpy()), "packed e2e obs mismatch"
    # headline: the host wire = ONE device-to-host copy of the observations in the palette format (the batch's
    # distinct cell values indexed with ceil(log2) bits per cell) + a 16-byte record per env (reward value and
    # counts, terminated mask, truncated); both checked for losslessness on the device, decoded by unpack_wire
    pal_bits, pal_codes = eng.wire_palette()
    e2e_pal_s, h = time_host("palette")
    e2e_s, h = time_host("wire")
    eng.check_status()  # (raises if a cell fell outside the palette or a reward did not fit its record)
    from multigrid_b200.engine import unpack_wire
    w_img, w_rew, w_term, w_trunc = unpack_wire(h["wire"], E, n, cfg["view"], pal_bits, pal_codes)
    assert np.array_equal(w_img[:M], eng.obs[:M].cpu().numpy()), "wire e2e obs mismatch"
    assert (w_rew == eng.reward.cpu().numpy()).all(), "wire e2e reward mismatch"
    assert np.array_equal(w_term, eng.terminated.cpu().numpy().astype(bool)), "wire e2e terminated mismatch"
    assert np.array_equal(w_trunc, eng.truncated.cpu().numpy().astype(bool)), "wire e2e truncated mismatch"
    checksum = int(h["wire"].sum(dtype=torch.int64)) + float(w_rew.sum())

    # ---- informational: the public Python env API, eager (no graph), device-resident actions
    api_env = envs[1]
    K3 = 2000
    for k in range(64):
        api_env.step(tape[k % n_tape])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K3):
        api_env.step(tape[k % n_tape])
    torch.cuda.synchronize()
    api_s = time.perf_counter() - t0
    clocks = sampler.stop()

    t = torch.tensor([ms, e2e_s, ms_closed, ms_chained or 0.0, clocks.get("sm_mhz") or 0.0, e2e_raw_s, e2e_p9_s, e2e_pal_s],
                     dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = torch.stack(allr).cpu().numpy()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, ms_closed, ms_chained_max, e2e_raw_s = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[5])
    e2e_p9_s, e2e_pal_s = float(t[6]), float(t[7])

    if rank == 0:
        W, H, V = cfg["W"], cfg["H"], cfg["view"]
        total_envs = E * world
        value = total_envs * n * K / (ms * 1e-3)
        bpe = algorithmic_bytes_per_env_step(W, H, n, V, cfg["mutable_grid"])
        moved = moved_bytes_per_env_step(W, H, n, eng.obs_stride, static)
        peak, peak_src = hbm_peak()
        us = 1e3 * ms / K
        achieved = bpe * E / us / 1e3  # GB/s per GPU: one launch = E env-steps
        grid_read = 3 * W * H
        traffic = ncu_traffic(args.config)
        bytes_io, bytes_raw = eng.bytes_per_step(packed="wire"), eng.bytes_per_step(packed=False)
        bytes_p9, bytes_pal = eng.bytes_per_step(packed=True), eng.bytes_per_step(packed="palette")
        kernel = ("mg::static_fast_kernel / static_rolled_kernel (MG_FLAG_STATIC_GRID: memoised views + agent overlay)"
                  if static else "mg::step_obs_kernel<V, MODE_STEP_OBS> (general kernel: per-env cells in shared memory)")
        line = {
            "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.config, world),
            "clocks": clocks,
            "e2e": {"value": total_envs * n * K2 / e2e_s, "unit": "agent-steps/s",
                    "h2d_bytes_per_step": bytes_io["h2d"], "d2h_bytes_per_step": bytes_io["d2h"],
                    "steps": K2, "timing": "wall clock around mg_step_obs_host_wire + stream sync per step",
                    "call": f"mg_step_obs_host_wire: ONE device-to-host copy = observations in the palette format "
                            f"({pal_bits} bits per cell = index into the {len(pal_codes)} cell values this batch can show) + "
                            f"{eng.lib.mg_wire_record_bytes(n)} bytes per env (reward value and per-agent counts, terminated "
                            "mask, truncated); lossless, checked on the device, decoded by engine.unpack_wire and "
                            "compared with the device tensors here",
                    "palette_bits": pal_bits, "palette_size": int(len(pal_codes)),
                    "palette": {"value": total_envs * n * K2 / e2e_pal_s, "d2h_bytes_per_step": bytes_pal["d2h"],
                                "pcie_gbs": (bytes_pal["h2d"] + bytes_pal["d2h"]) * K2 / e2e_pal_s / 1e9 / world,
                                "call": "mg_step_obs_host_palette (palette observations; rewards f64 x n, terminated, "
                                        "truncated as four copies)"},
                    "pcie_gbs": (bytes_io["h2d"] + bytes_io["d2h"]) * K2 / e2e_s / 1e9 / world,
                    "pcie_frac_of_64GBs": (bytes_io["h2d"] + bytes_io["d2h"]) * K2 / e2e_s / 1e9 / world / 64.0,
                    "packed9": {"value": total_envs * n * K2 / e2e_p9_s, "d2h_bytes_per_step": bytes_p9["d2h"],
                                "pcie_gbs": (bytes_p9["h2d"] + bytes_p9["d2h"]) * K2 / e2e_p9_s / 1e9 / world,
                                "call": "mg_step_obs_host_packed (9 bits per cell, no palette)"},
                    "unpacked": {"value": total_envs * n * K2 / e2e_raw_s, "d2h_bytes_per_step": bytes_raw["d2h"],
                                 "pcie_gbs": (bytes_raw["h2d"] + bytes_raw["d2h"]) * K2 / e2e_raw_s / 1e9 / world,
                                 "call": "mg_step_obs_host (raw 3-byte cells)"},
                    "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] if len(os.sched_getaffinity(0)) > 4
                                    else sorted(os.sched_getaffinity(0)),
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                "traffic_source": None if traffic is None else traffic.get("source"),
                "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0,
                "plain_us": us, "plain_frac": achieved / peak,
                "launches": "plain: `value`, `achieved` and `frac` are all measured with plain launches",
                "algorithmic_bytes_per_env_step": bpe, "algorithmic_bytes_per_launch": bpe * E,
                # the static-grid kernel (and the single-layout dedup before it) does not read the env's grid from
                # HBM at all: the same launch time against the algorithmic bytes WITHOUT that read
                "frac_excluding_dedup": ((bpe - grid_read) if static else bpe) * E / us / 1e3 / peak,
                "moved_bytes_per_env_step": moved, "moved_gbs": moved * E / us / 1e3,
                "frac_moved": moved * E / us / 1e3 / peak,
                "kernel": kernel, "static_grid_path": static, "obs_agent_stride": eng.obs_stride,
            },
            "closed_loop": {
                "note": "informational: the same K plain launches on ONE batch stepped back to back (its state and "
                        "outputs stay in L2; what a closed-loop caller with a device-side policy sees)",
                "us_per_launch": 1e3 * ms_closed / K, "value": total_envs * n * K / (ms_closed * 1e-3)},
            "api_device_resident": {
                "note": "informational: env.step(actions_on_device) of the public Python API, eager launches, per-agent "
                        "dict results left on the device, one batch stepped back to back, this rank",
                "us_per_step": 1e6 * api_s / K3, "value": E * n * K3 / api_s, "steps": K3},
            "verified": verified,
        }
        if ms_chained is not None:
            line["chained"] = {
                "note": "informational: the same K launches with MG_FLAG_CHAINED (per-env tickets instead of the "
                        "kernel-boundary barrier; pays only across DIFFERENT batches, as here)",
                "us_per_launch": 1e3 * ms_chained_max / K, "value": total_envs * n * K / (ms_chained_max * 1e-3),
                "achieved_gbs": bpe * E / (1e3 * ms_chained_max / K) / 1e3}
        if per_rank is not None:
            line["per_rank"] = {"graph_ms": [float(v) for v in per_rank[:, 0]], "e2e_s": [float(v) for v in per_rank[:, 1]],
                                "sm_mhz": [float(v) for v in per_rank[:, 4]],
                                "note": "value uses the slowest rank's graph time (max over ranks)"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="empty8", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
