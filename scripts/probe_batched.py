"""Scratch GPU probe: stand-alone sweep time for m = 1..64 pending terms and the step time at m lines per scan."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ms_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 32, 64]
steps = 30
n = 3 + 2 * N
f = EkfFilter(capacity_lines=N + 512)
scn0 = sc.map_scenario(N, 1, m=8, seed=1)
f.scan(np.zeros(3), scn0["seed_z"], scn0["seed_R"])
for mm in (1, 8, 12, 16, 24, 32, 48, 64):
    ms = f.sweep_probe(m=mm, repeats=5)
    b = 8.0 * n * (n + 1) + 32.0 * n * mm
    print("sweep m=%2d  %.3f ms  %.1f GB/s algorithmic  %.2f TFLOP/s fp64" % (mm, ms, b / ms / 1e6, 4.0 * mm * n * (n + 1) / 2 / ms / 1e9), flush=True)
f.close()
for m in ms_list:
    scn = sc.map_scenario(N, steps + 3, m=m, seed=1, stride=m + 3)
    f = EkfFilter(capacity_lines=N + 512)
    f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    for s in range(3):
        f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
    f.sync()
    f.profile_read(); f.profile_enable(True)
    t = time.time(); nm = 0
    for s in range(3, steps + 3):
        rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        nm += int((j >= 0).sum())
    f.sync()
    dt = time.time() - t
    pr = f.profile_read()
    print("m=%2d  %.3f ms/step (e2e)  matches %d/%d  sweep launches %d  sweep_ms/step %.3f  launches/step %.1f" %
          (m, 1e3 * dt / steps, nm, steps * m, pr["sweeps"], pr["sweep_ms"] / steps, pr["launches"] / steps), flush=True)
    f.close()
