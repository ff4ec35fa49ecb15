"""Minimal driver for ncu captures of the covariance sweep: seeds an N-landmark map and runs the
stand-alone sweep (ekf_sweep_probe) for the pending-term counts given on the command line."""
import sys

import numpy as np

sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ms_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8]
scn = sc.map_scenario(N, 1, m=8, seed=1)
f = EkfFilter(capacity_lines=N + 256)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
n = 3 + 2 * N
for mm in ms_list:
    ms = f.sweep_probe(m=mm, repeats=1)
    print("m=%d %.3f ms %.1f GB/s" % (mm, ms, (8.0 * n * (n + 1) + 32.0 * n * mm) / ms / 1e6))
