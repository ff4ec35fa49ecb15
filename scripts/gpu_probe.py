"""Scratch GPU probe: sweep bandwidth and step time at a given map size (not part of the test-suite)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
m = 8
scn = sc.map_scenario(N, steps, m=m, seed=1)
import os
flags = int(os.environ.get("EKF_FLAGS", "0"))
f = EkfFilter(capacity_lines=N + 256, flags=flags)
t = time.time()
rc, j, pose = f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
print("seed rc", rc, "L", f.lines, "%.3fs" % (time.time() - t), flush=True)
n = 3 + 2 * N
for mm in (1, 2, 4, 8, 16, 32):
    ms = f.sweep_probe(m=mm, repeats=5)
    b = 8.0 * n * (n + 1) + 32.0 * n * mm
    print("sweep m=%2d  %.3f ms  %.1f GB/s algorithmic" % (mm, ms, b / ms / 1e6), flush=True)
f.profile_enable(True)
t = time.time()
nm = 0
for s in range(steps):
    rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
    nm += int((j >= 0).sum())
dt = time.time() - t
pr = f.profile_read()
print("steps %d  %.3f ms/step (host wall, e2e)  matches %d/%d  sweeps %d  sweep_ms/step %.3f  launches/step %.1f" %
      (steps, 1e3 * dt / steps, nm, steps * m, pr["sweeps"], pr["sweep_ms"] / max(pr["sweeps"], 1), pr["launches"] / steps))
print("sweep GB/s in-step: %.1f" % (pr["sweep_bytes"] / pr["sweep_ms"] / 1e6))
