"""Scratch: a few m = 64 scans at N landmarks (used under compute-sanitizer)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc  # noqa: E402
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
m = 64
scn = sc.map_scenario(N, 6, m=m, seed=1, stride=m + 3)
f = EkfFilter(capacity_lines=N + 512)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
for s in range(6):
    rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
    print(s, rc, int((j >= 0).sum()), flush=True)
f.sync()
print("ok", f.lines)
