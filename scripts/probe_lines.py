"""Scratch: device line extraction vs the CPU oracle / literal reference -- agreement and time per scan."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from slam_ros_b200 import LineExtractor, scenario as sc  # noqa: E402
from oracle.oracle import LinesOracle  # noqa: E402

S = sc.room_scans(steps=50, seed=17, range_sigma=2e-3)
lx = LineExtractor(); lo = LinesOracle()
rows, n = lx.extract(S["scans"][0]); ref, m = lo.extract(S["scans"][0])
print("lines", n, m)
if n == m and n:
    print("max |d alfa| %.2e  |d r| %.2e  C rel %.2e" % (np.abs(rows[:, 0] - ref[:, 0]).max(), np.abs(rows[:, 1] - ref[:, 1]).max(),
                                                     (np.abs(rows[:, [2, 5]] - ref[:, [2, 5]]) / ref[:, [2, 5]]).max()))
t = time.perf_counter()
for s in range(50):
    lx.extract(S["scans"][s])
dt = (time.perf_counter() - t) / 50
t = time.perf_counter()
for s in range(10):
    lo.extract(S["scans"][s])
dc = (time.perf_counter() - t) / 10
print("device %.1f us per scan (host payload in, lines out), CPU restatement %.1f us" % (dt * 1e6, dc * 1e6))
