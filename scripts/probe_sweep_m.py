"""Scratch: stand-alone sweep for every pending-term count 1..64 (finds a failing count)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc  # noqa: E402
N = 3000
scn = sc.map_scenario(N, 1, m=8, seed=1)
f = EkfFilter(capacity_lines=N + 512)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
for mm in range(1, 65):
    try:
        ms = f.sweep_probe(m=mm, repeats=1)
    except Exception as e:
        print("m=%d FAILED: %s" % (mm, e)); break
else:
    print("all 64 counts ok")
