"""Row-sharded P over >= 2 GPUs: the NCCL exchange of the H-column slices and the fused in-kernel NVLink
exchange (CUDA IPC peer memory), the latter also at a size where the sweep overlaps the next scan's line loop.
Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(world, n_landmarks, steps, mode, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_gpu", "sharded_check.py"), str(n_landmarks), str(steps), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sharded x%d OK (%s)" % (world, mode) in out.stdout


@pytest.mark.parametrize("mode,n_landmarks,steps,port", [
    ("nccl", 300, 25, 29517),
    ("fused", 300, 25, 29518),      # synchronous path: cooperative line loop, sweep after it
    ("fused", 3300, 12, 29519),     # n = 6603 >= the overlap threshold: sweep of scan s under the line loop of scan s+1
    ("switch", 3300, 12, 29520),    # alternates between the NCCL and the fused exchange every 3 scans (same bits either way)
])
def test_row_sharded_filter_two_ranks(libekf, mode, n_landmarks, steps, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(2, n_landmarks, steps, mode, port)


def test_ipc_handle_needs_a_sharded_filter(libekf):
    from slam_ros_b200 import EkfFilter
    from slam_ros_b200.ekf import EkfError, EKF_ESTATE
    f = EkfFilter(capacity_lines=64, device=0)
    with pytest.raises(EkfError) as e:
        f.shard_ipc_handle()
    assert e.value.code == EKF_ESTATE
