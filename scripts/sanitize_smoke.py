"""Small run that launches every kernel of libekfcuda once or a few times -- the subject of the compute-sanitizer passes
(memcheck / racecheck / synccheck, one tool per gpurun call; logs kept under profiles/):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_ros_b200 import EkfFilter, EkfBatch, LineExtractor, scenario as sc  # noqa: E402

big = "--no-overlap-case" not in sys.argv
# 1. small map: cluster line loop, in-place sweeps (DFMA consumers at 8 terms, tensor-core consumers at 16), step-wise kernels
N = 300
scn = sc.map_scenario(N, 6, m=8, seed=2)
f = EkfFilter(capacity_lines=N + 64)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
for s in range(2):
    rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
    assert rc == 0 and (j >= 0).all()
scn16 = sc.map_scenario(N, 2, m=16, seed=2, stride=17)
rc, j, pose = f.scan(scn16["u"][0], scn16["z"][0], scn16["R"][0])
f.predict(scn["u"][3])
for i in range(8):
    jj, innov = f.associate(scn["z"][3][i], scn["R"][3][i])
    if jj >= 0:
        f.update(jj, scn["z"][3][i], scn["R"][3][i])
    else:
        f.add_line(scn["z"][3][i], scn["R"][3][i])
f.end_scan(8)
z_far = scn["z"][4].copy(); z_far[:, 1] += 3.0                 # nothing matches: augmentation kernels
f.scan(scn["u"][4], z_far, scn["R"][4])
y, P, L = f.download_live(); f.cov_stats(); f.get_ellipse(); f.download_block(3, 3, 40, 40)
f.sweep_probe(m=3, repeats=1)
f.close()
print("single filter, small map: ok", flush=True)
# 2. overlapped path (n >= 6000): cooperative line loop beside the out-of-place sweep on the second stream
if big:
    N = 3100
    scn = sc.map_scenario(N, 4, m=8, seed=3)
    f = EkfFilter(capacity_lines=N + 64)
    f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    for s in range(3):
        rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        assert rc == 0
    f.download_block(0, 0, 16, 16)
    f.close()
    print("single filter, overlapped path: ok", flush=True)
# 3. Monte-Carlo batch: on-chip triangle, augmentation, and (EKF_BATCH_NS) the off-chip fallback
for ns in (None, "51"):
    if ns:
        os.environ["EKF_BATCH_NS"] = ns
    B, N = 6, 50
    scns = [sc.map_scenario(N, 4, m=8, seed=40 + k) for k in range(B)]
    bt = EkfBatch(B, capacity_lines=64)
    os.environ.pop("EKF_BATCH_NS", None)
    bt.scan(np.zeros((B, 3)), np.stack([x["seed_z"] for x in scns]), np.stack([x["seed_R"] for x in scns]))
    for s in range(3):
        Z = np.stack([x["z"][s] for x in scns])
        if s == 1:
            Z = Z.copy(); Z[:, :, 1] += 3.0
        bt.scan(np.stack([x["u"][s] for x in scns]), Z, np.stack([x["R"][s] for x in scns]))
    bt.download(2)
    bt.close()
print("batch: ok", flush=True)
# 4. line extraction
lx = LineExtractor()
S = sc.room_scans(steps=2, seed=4)
for s in range(2):
    rows, n = lx.extract(S["scans"][s])
    assert n > 5
print("line extraction: ok", flush=True)
