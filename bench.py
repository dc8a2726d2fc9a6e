#!/usr/bin/env python
"""Benchmark of the MultiGrid step/observe hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU oracle port, all host threads
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, env axis sharded

Workload (BASELINE.json configs[1], per GPU): MultiGrid-Empty-8x8-v0, agents=4, view 7,
num_envs=65536, uniform random actions over the 7 actions, "next-step" auto-reset; a step is one
fused mg_step_obs launch over the whole batch. Metric: agent-steps/s (1 agent-step = one agent
slot of one env advanced by one step; terminated/skipped agents count, on CPU and GPU alike).

One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent-steps/sec on Empty-8x8 agents=4 num_envs=65536; HBM GB/s vs 8 TB/s peak"
SIZE, N_AGENTS, VIEW, ENVS_PER_GPU = 8, 4, 7, 65536
MAX_STEPS = 4 * SIZE * SIZE  # envs/empty.py:145
REPLICAS = 8   # state replicas rotated through so each launch finds its inputs in HBM, not L2
BURN_IN = 64   # untimed steps that de-synchronise the envs before anything is measured
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used if MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum of one steady-state launch of the fused kernel, from the
# `ncu --set full` capture summarised in profiles/r01_summary.md (r01c): 26.30 MB read (= the layout's
# 401 B/env of inputs) + 3.04 MB written; the other ~41.6 MB of outputs are still dirty in the 126 MB L2
# when the kernel ends (ncu flushes before each replay) and reach HBM during later launches.
NCU_DRAM_BYTES_PER_LAUNCH = 7608576  # profiles/r01j_ncu_details.txt: dram read 6 120 704 + write 1 487 872 (single-layout dedup)


def algorithmic_bytes_per_env_step(W, H, n, V, mutable_grid=False):
    """packed state round trip + outputs, per env-step."""
    reads = 3 * W * H + 7 * n + 2 + 16 + 16 + n
    writes = (3 * W * H if mutable_grid else 0) + 7 * n + 2 + 16 + 3 * n * V * V + 8 * n + n + 1
    return reads + writes


def rollout_bytes_per_env_step(n, V):
    """SURVEY.md §8(d), in-kernel multi-step rollout: state stays on chip, so the per-env-step floor
    is actions in + outputs out (obs, f64 reward, terminated, truncated, direction)."""
    return n + 3 * n * V * V + 8 * n + n + 1 + n


def layout_bytes_per_env_step(W, H, n, V, auto_reset=True):
    """Bytes the engine's HBM layout actually moves per env-step: padded 4-byte
    cell words, 8-byte agent records, int32 counters, 148-byte obs slots."""
    ostride = (3 * V * V + 3) & ~3
    reads = 4 * (W + 1) * (H + 1) + 8 * n + 4 + 16 + 16 + n + (4 if auto_reset else 0)
    writes = 8 * n + 4 + 16 + n * ostride + 8 * n + n + 1 + (4 if auto_reset else 0)
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


# -------------------------------------------------------------------------------------------------
# CPU side: the oracle port (oracle/mg_oracle.c) -- bench.py's cpu_baseline / --impl reference leg
# -------------------------------------------------------------------------------------------------
def make_cpu_oracle(num_envs, nthreads):
    from oracle.c_oracle import COracle, build
    from oracle.mg_oracle import OracleConfig
    build()
    cfg = OracleConfig(W=SIZE, H=SIZE, n=N_AGENTS, V=VIEW, max_steps=MAX_STEPS, auto_reset=True)
    pg, pa = empty_layout(SIZE, N_AGENTS)
    st, inc = pcg_words(0, num_envs)
    return COracle(cfg, np.repeat(pg, num_envs, 0), np.repeat(pa, num_envs, 0), st, inc,
                   pool_grid=pg, pool_agents=pa, nthreads=nthreads)


def time_cpu_oracle(num_envs, steps, warmup, nthreads, budget_s=None):
    """Returns (agent_steps_per_s, steps_done, seconds)."""
    ora = make_cpu_oracle(num_envs, nthreads)
    rng = np.random.default_rng(0)
    tape = rng.integers(0, 7, size=(64, num_envs, N_AGENTS)).astype(np.int8)
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
    return num_envs * N_AGENTS * done / dt, done, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once per box, on rank 0
    cores = len(os.sched_getaffinity(0))
    value, done, dt = time_cpu_oracle(ENVS_PER_GPU, args.steps, args.warmup, cores)
    sample = f"{done} steps x {ENVS_PER_GPU} envs ({dt:.2f} s)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": workload_config(args.gpus) | {"note": "CPU oracle port (oracle/mg_oracle.c, OpenMP); the reference itself is Python+numba and cannot travel to the GPU box"},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {
        "workload": "MultiGrid-Empty-8x8-v0 agents=4 view=7 num_envs=65536 per GPU, uniform random "
                    "actions, next-step auto-reset (BASELINE.json configs[1]; configs[4] at 8 GPUs)",
        "num_envs_per_gpu": ENVS_PER_GPU, "num_envs_total": ENVS_PER_GPU * n_gpus,
        "agents": N_AGENTS, "view_size": VIEW, "grid": f"{SIZE}x{SIZE}", "max_steps": MAX_STEPS,
        "sharding": f"env axis split over {n_gpus} GPU(s), no collective on the step path",
        "l2": f"{REPLICAS} state replicas rotated (each launch touches a batch last used "
              f"{REPLICAS} launches ago; {REPLICAS}x61 MB > 126 MB L2), no explicit flush; "
              "MG_FLAG_STREAM_STATE (L2 evict_first on state loads / obs stores, cache policy only)",
        "dedup": "single-layout dedup (MgState.grid_dirty / pool_rep): Empty-8x8 has ONE reset layout, so groups whose envs "
                 "still equal it take their cells from an L2-resident 32-copy buffer instead of reading their 324-byte grid "
                 "copies from HBM (the algorithmic 929 B per env-step still count the 192-byte grid read)",
        "launches": "chained (MG_FLAG_CHAINED): each launch is ordered after the previous launch on the same replica env by "
                    "env through chain tickets, not by a kernel-boundary barrier, so it loads while its predecessor drains; "
                    "`unchained` in this line is the same graph with plain launches",
    }


# -------------------------------------------------------------------------------------------------
# GPU side
# -------------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist
    from multigrid_b200 import _cabi
    from multigrid_b200.engine import EngineConfig, StepEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _cabi.load()
    K, Wm = args.steps, max(args.warmup, 3)
    E, n = ENVS_PER_GPU, N_AGENTS

    cfg = EngineConfig(width=SIZE, height=SIZE, num_agents=n, view_size=VIEW, max_steps=MAX_STEPS,
                       auto_reset=True, stream_state=True)
    pg, pa = empty_layout(SIZE, n)
    engines = []
    for r in range(REPLICAS):
        eng = StepEngine(cfg, E, dev, pg, pa)
        first = (rank * REPLICAS + r) * E  # seeds are a function of the global env id
        st, inc = pcg_words(first, E)
        eng.load_state(pcg_state=st, pcg_inc=inc)
        eng.reset_from_pool()  # every env starts from the (single) pool layout, like env.reset()
        engines.append(eng)

    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    n_tape = max(64, K)  # no action set is replayed inside the timed region (a short cycle keeps agents
                         # near their start cells, which flatters the kernel by ~15 %)
    tape = torch.randint(0, 7, (n_tape, E, n), generator=gen, device=dev, dtype=torch.int32).to(torch.int8)

    def launch(k, chained=True):
        # MG_FLAG_CHAINED: consecutive launches are ordered per env by the engine's chain tickets instead of a
        # kernel-boundary barrier (the action tape is staged before the timed region, nothing else runs on the stream)
        engines[k % REPLICAS].step(tape[k % n_tape], chained=chained)

    for k in range(BURN_IN):
        launch(k)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- `value`: K fused launches, inputs resident in HBM, captured in one CUDA graph --------
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for k in range(Wm):
            launch(k)
        stream.synchronize()
        before = lib.mg_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(K):
                launch(Wm + k)
        launches = lib.mg_launch_count() - before
        graph.replay()  # untimed: the first replay of an instantiated graph also uploads it to the device
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        graph.replay()
        ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)

    # ---- informational: the same K launches as PLAIN launches (every launch waits for the whole previous grid)
    with torch.cuda.stream(stream):
        graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph2, stream=stream):
            for k in range(K):
                launch(Wm + k, chained=False)
        graph2.replay()
    torch.cuda.synchronize()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev2.record(stream)
        graph2.replay()
        ev3.record(stream)
    torch.cuda.synchronize()
    ms2 = ev2.elapsed_time(ev3)

    # ---- `e2e`: the host-buffer call (mg_step_obs_host): H2D actions, kernel, D2H results -------
    eng = engines[0]
    h = eng.host_buffers()
    host_tape = tape[:8].cpu()
    K2 = max(3, min(K, 50))
    for k in range(3):
        h["actions"].copy_(host_tape[k % 8])
        eng.step_host()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(K2):
        h["actions"].copy_(host_tape[k % 8])  # the caller's actions land in the pinned buffer
        eng.step_host(synchronize=True)       # results are in pinned host memory on return
    e2e_s = time.perf_counter() - t0
    checksum = int(h["obs"].view(torch.uint8).sum()) + float(h["reward"].sum())

    # ---- informational: the public Python env API, eager (no graph), device-resident actions: one
    # BatchedMultiGridEnv stepped back to back (its 61 MB of state + outputs stay L2-resident)
    from multigrid_b200.envs import make
    api_env = make("MultiGrid-Empty-8x8-v0", agents=n, num_envs=E, device=dev, auto_reset=True, first_env=rank * E)
    api_env.reset(seed=1234)
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

    t = torch.tensor([ms, e2e_s, ms2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, ms2 = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        total_envs = E * world
        value = total_envs * n * K / (ms * 1e-3)
        bpe = algorithmic_bytes_per_env_step(SIZE, SIZE, n, VIEW)
        peak, peak_src = hbm_peak()
        achieved = bpe * E / (ms * 1e-3 / K) / 1e9  # per GPU: one launch = E envs
        bytes_io = eng.bytes_per_step()
        line = {
            "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": total_envs * n * K2 / e2e_s, "unit": "agent-steps/s",
                    "h2d_bytes_per_step": bytes_io["h2d"], "d2h_bytes_per_step": bytes_io["d2h"],
                    "steps": K2, "timing": "wall clock around mg_step_obs_host + stream sync per step",
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_note": "ncu dram read+write bytes of one launch (profiles/r01j_ncu_details.txt); "
                                         "outputs still dirty in L2 at kernel end are not in it; the grids come from the L2-resident layout buffer (dedup)",
                         "algorithmic_bytes_per_launch": bpe * E, "peak_source": peak_src,
                         "frac_of_8TBs_nominal": achieved / 8000.0,
                         "algorithmic_bytes_per_env_step": bpe,
                         "actual_bytes_per_env_step": layout_bytes_per_env_step(SIZE, SIZE, n, VIEW),
                         "kernel": "mg::step_obs_kernel<7, MODE_STEP_OBS, MULTI=false, CHAIN=true> (one launch = 65536 env-steps)",
                         "launch_time_note": "avg_launch_us = timed region / launches; chained launches overlap (a launch "
                                             "loads while its predecessor drains), so this is the throughput time per "
                                             "launch, not the latency of one launch (see `unchained`)",
                         "avg_launch_us": 1e3 * ms / K},
        }
        line["unchained"] = {
            "note": "informational: the same K launches as plain launches (kernel-boundary barrier between them)",
            "us_per_launch": 1e3 * ms2 / K, "value": total_envs * n * K / (ms2 * 1e-3),
            "achieved_gbs": bpe * E / (ms2 * 1e-3 / K) / 1e9}
        line["api_device_resident"] = {
            "note": "informational: env.step(actions_on_device) of the public Python API, eager launches, per-agent "
                    "dict results left on the device, one batch stepped back to back (state stays in L2), this rank",
            "us_per_step": 1e6 * api_s / K3, "value": E * n * K3 / api_s, "steps": K3}
        if world == 1 and not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            v, done, dt = time_cpu_oracle(E, 10**9, 2, cores, budget_s=args.cpu_seconds)
            v1, done1, dt1 = time_cpu_oracle(4096, 10**9, 2, 1, budget_s=3.0)
            line["cpu_baseline"] = {
                "value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                "sample": f"{done} steps x {E} envs of the same workload ({dt:.1f} s), C oracle "
                          f"port with OpenMP; 1 thread: {v1:.3g} agent-steps/s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
