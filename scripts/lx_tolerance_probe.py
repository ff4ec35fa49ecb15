import sys, os
sys.path.insert(0, '.')
import numpy as np
from slam_ros_b200 import LineExtractor, scenario as sc
from oracle.oracle import LinesOracle
lo = LinesOracle(); lx = LineExtractor(max_lines=64)
worst = 0.0; worst_abs=0; cnt=0
g = np.load('tests/golden/lines_literal.npz')
def upd(rows, ref):
    global worst, cnt
    for c in (2,5):
        rel = np.abs(rows[:, c] - ref[:, c]) / np.maximum(np.abs(ref[:, c]), 1e-300)
        if rel.size: worst = max(worst, rel.max()); cnt += rel.size
for s in range(g["scans"].shape[0]):
    rows, n = lx.extract(g["scans"][s]); m = int(g["count"][s]); upd(rows[:m], g["rows"][s,:m])
print('golden worst', worst, cnt)
for seed, sig in ((17,2e-3),(5,1e-3),(9,5e-3),(23,1e-2)):
    S = sc.room_scans(steps=60, seed=seed, range_sigma=sig)
    for s in range(60):
        rows, n = lx.extract(S["scans"][s]); ref, m = lo.extract(S["scans"][s])
        assert n == m
        upd(rows, ref)
    print('room seed', seed, 'sigma', sig, 'worst so far', worst, cnt)
