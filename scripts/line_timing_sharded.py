"""Phase timestamps of the row-sharded line-loop kernel (library built with -DEKF_LINE_TIMING, EKF_LIB=...).
torchrun --nproc-per-node G scripts/line_timing_sharded.py [N]"""
import ctypes as C
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc, load_library  # noqa: E402
from slam_ros_b200.ekf import nccl_unique_id  # noqa: E402
from slam_ros_b200.parallel import connect_shards  # noqa: E402
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device=dev)
dist.broadcast(uid, 0)
scn = sc.map_scenario(N, 6, m=8, seed=1)
f = EkfFilter(capacity_lines=N + 256, device=local, shard=(rank, world, bytes(uid.cpu().tolist())))
assert connect_shards(f, dev)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
for s in range(6):
    f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
f.sync()
lib = load_library()
buf = (C.c_ulonglong * (16 * 16))()
rc = lib.ekf_debug_line_timing(buf, 16 * 16)
t = np.array(list(buf), dtype=np.int64).reshape(16, 16)
names = ["start", "landmarks", "blockmin", "barrier1", "pre-rows", "staged", "rows/part2", "barrier2", "part1", "fence+sync", "xchg", "sync2"]
order = [0, 1, 2, 3, 4, 8, 9, 10, 11, 6, 7]
import time
time.sleep(0.5 * rank)
if True:
    for line in range(8):
        row = t[line]
        if row[0] == 0:
            continue
        base = row[0]
        print("rank", rank, "line", line, "t0=%.1f" % ((base - t[0, 0]) / 1e3), " ".join("%s=%+.1f" % (names[i], (row[i] - base) / 1e3) for i in order if row[i] > 0),
              "| next %+.1f" % ((t[line + 1, 0] - base) / 1e3) if t[line + 1, 0] > 0 else "", flush=True)
dist.barrier()
dist.destroy_process_group()
