"""Scratch: device line extraction vs the CPU oracle on messy payloads (outliers, duplicated beams, dropped sectors,
random order).  Reports any payload where the line count or a line differs beyond the test tolerances."""
import sys
import numpy as np
sys.path.insert(0, ".")
from slam_ros_b200 import LineExtractor, scenario as sc  # noqa: E402
from oracle.oracle import LinesOracle  # noqa: E402

lx = LineExtractor(); lo = LinesOracle()
rng = np.random.default_rng(123)
S = sc.room_scans(steps=40, seed=99, range_sigma=3e-3)["scans"]
bad = 0
for t in range(120):
    p = S[t % 40].copy()
    kind = t % 6
    if kind == 0:                                   # 5 % outliers
        k = rng.random(p.shape[0]) < 0.05; p[k, 0] = rng.uniform(0.1, 9.0, k.sum())
    elif kind == 1:                                 # duplicated beams
        idx = rng.integers(0, p.shape[0], 30); p = np.concatenate([p, p[idx]])
    elif kind == 2:                                 # dropped sectors
        a = rng.integers(0, 300); p[a:a + rng.integers(5, 60), 0] = 0.0
    elif kind == 3:                                 # random order
        p = p[rng.permutation(p.shape[0])]
    elif kind == 4:                                 # heavy noise
        p[:, 0] += (rng.standard_normal(p.shape[0]) * 0.03).astype(np.float32) * (p[:, 0] > 0)
    else:                                           # sparse
        p = p[::rng.integers(2, 6)]
    rows, n = lx.extract(p); ref, m = lo.extract(p)
    ok = (n == m)
    if ok and n:
        d = np.abs(rows[:, 0] - ref[:, 0]); d = np.minimum(d, np.abs(d - 2 * np.pi))
        ok = d.max() < 1e-9 and np.abs(rows[:, 1] - ref[:, 1]).max() < 1e-9
        ok = ok and (np.abs(rows[:, [2, 5]] - ref[:, [2, 5]]) / np.maximum(np.abs(ref[:, [2, 5]]), 1e-300)).max() < 1e-3
    if not ok:
        bad += 1
        print("payload %d (kind %d): %d vs %d lines" % (t, kind, n, m))
print("fuzz: %d of 120 payloads differ" % bad)
