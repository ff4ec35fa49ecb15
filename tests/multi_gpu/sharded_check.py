"""Row-sharded covariance across ranks (one process per GPU) against the CPU oracle.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port 29511 tests/multi_gpu/sharded_check.py [N_landmarks] [steps]
Every rank replays the same scan sequence on its shard; association indices must be bit-exact on every rank
and the ranks' partial read-outs must sum to the oracle's covariance within 1e-9."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slam_ros_b200 import EkfFilter, scenario as sc  # noqa: E402
from slam_ros_b200.ekf import nccl_unique_id  # noqa: E402
from oracle.oracle import StructuredOracle  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, 0)
    cap = N + 64
    f = EkfFilter(capacity_lines=cap, device=local, shard=(rank, world, bytes(uid.cpu().tolist())))
    mode = sys.argv[3] if len(sys.argv) > 3 else "fused"
    if mode in ("fused", "switch"):         # in-kernel NVLink exchange; "nccl" keeps the ncclAllReduce path
        from slam_ros_b200.parallel import connect_shards
        assert connect_shards(f, dev), "CUDA IPC peer mapping failed"
    so = StructuredOracle(cap)
    scn = sc.map_scenario(N, steps, m=8, seed=5)
    rc, j, pose = f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    so.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    assert rc == 0 and f.lines == so.lines == N
    for s in range(steps):
        if mode == "switch" and s % 3 == 0:     # every rank leaves / re-enters the fused exchange at the same step
            f.shard_use_fused((s // 3) % 2 == 1)
        rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        st, jo = so.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        assert np.array_equal(j, jo), "rank %d step %d: %s vs %s" % (rank, s, j, jo)
        assert np.abs(pose - so.pose).max() < 1e-10
    y, Ppart, L = f.download_live()
    yo, Po = so.live()
    assert L == so.lines
    assert np.abs(y - yo).max() / np.abs(yo).max() < 1e-9           # y is replicated
    t = torch.tensor(Ppart, dtype=torch.float64, device=dev)
    dist.all_reduce(t)                                               # partial read-outs sum to the full matrix
    P = t.cpu().numpy()
    err = np.abs(P - Po).max() / np.abs(Po).max()
    assert err < 1e-9, err
    tr, sm, sq = f.cov_stats()
    tt = torch.tensor([tr, sm, sq], dtype=torch.float64, device=dev)
    dist.all_reduce(tt)
    assert abs(float(tt[0]) - np.trace(Po)) / abs(np.trace(Po)) < 1e-9
    ms = f.sweep_probe(m=8, repeats=3)
    if rank == 0:
        print("sharded x%d OK (%s): N=%d steps=%d  P rel err %.2e  local sweep %.3f ms" % (world, mode, N, steps, err, ms), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
