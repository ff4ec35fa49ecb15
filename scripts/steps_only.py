"""Seeds an N-landmark map and runs a few scans (for ncu launch lists)."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
flags = int(os.environ.get("EKF_FLAGS", "0"))
scn = sc.map_scenario(N, steps, m=8, seed=1)
f = EkfFilter(capacity_lines=N + 256, flags=flags)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
for s in range(steps):
    rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
print("ok", f.lines, j.tolist())
