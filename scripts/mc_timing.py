"""Phase cycle totals of one Monte-Carlo filter's scan (library built with -DEKFB_TIMING, loaded through EKF_LIB):
    EKF_LIB=slam_ros_b200/libekfcuda_mctiming.so python scripts/mc_timing.py [filters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slam_ros_b200 import EkfBatch, scenario as sc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, m, cap = 50, 8, 64
scns = [sc.map_scenario(N, 4, m=m, seed=1000 + k) for k in range(64)]
reps = (B + 63) // 64
tile = lambda key: np.concatenate([np.stack([x[key] for x in scns])] * reps)[:B]
bt = EkfBatch(B, capacity_lines=cap, device=0)
bt.scan(np.zeros((B, 3)), tile("seed_z"), tile("seed_R"))
for s in range(3):
    print("--- scan %d (%d filters)" % (s, B), flush=True)
    bt.scan(np.stack([x["u"][s] for x in scns] * reps)[:B], np.concatenate([np.stack([x["z"][s] for x in scns])] * reps)[:B],
            np.concatenate([np.stack([x["R"][s] for x in scns])] * reps)[:B])
