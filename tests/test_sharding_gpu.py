"""Hardware shard-equivalence check: a global env batch split over 2 and 4 ranks, launched with
torch.distributed.run like the bench, gives the same per-GLOBAL-env checksums of (obs, reward, terminated,
truncated, final state) as one rank stepping the whole batch -- on the real CUDA engine, for the static-grid kernel
(Empty-8x8) and for the general kernel with a device-generated layout pool (BlockedUnlockPickup). The ranks use one
GPU each when the box has them and share cuda:0 otherwise."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, env_id, agents, total, T, out, auto_reset):
    cmd = [sys.executable]
    if world > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                "--master-port", str(_free_port())]
    cmd += [os.path.join(ROOT, "tests", "shard_worker.py"), env_id, str(agents), str(total), str(T), out,
            "1" if auto_reset else "0"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    return np.load(out)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("env_id,agents,total,auto_reset", [
    ("MultiGrid-Empty-8x8-v0", 4, 4100, True),             # static-grid kernel, ragged shards, auto-reset at 40 steps
    ("MultiGrid-BlockedUnlockPickup-v0", 2, 1030, False),  # general kernel, per-env layouts from the device generators
])
def test_shards_equal_the_single_rank_run(env_id, agents, total, auto_reset, tmp_path):
    T = 60
    one = _run(1, env_id, agents, total, T, str(tmp_path / "w1.npy"), auto_reset)
    assert len(np.unique(one)) > total // 2  # the checksums do tell envs apart
    for world in (2, 4):
        got = _run(world, env_id, agents, total, T, str(tmp_path / f"w{world}.npy"), auto_reset)
        bad = np.nonzero(got != one)[0]
        assert bad.size == 0, f"world {world}: {bad.size} global envs differ, first {bad[:8]}"
