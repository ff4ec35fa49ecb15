"""Golden fingerprints of the CPU oracle at BASELINE.json's FULL config sizes -- TEST INFRASTRUCTURE.

    python tests/golden/make_golden_fullsize.py 3300 | 1k | 10k | 40k | all   (comma-separated lists work too)

The structured oracle (oracle/ekf_oracle.cpp, bitwise equal to the literal reference where the literal reference can
run: tests/test_oracle.py) is replayed here, on the CPU, over the same seeded scan sequences bench.py uses
(scenario.map_scenario(N, steps, m = 8, seed = 1)); what is kept per config is small enough to commit:

    j_out[steps, m]   the association of every line of every step           (compared bit for bit)
    pose[steps, 3]    xPos, yPos, thetaPos after every step
    at every `every`-th step and at the last one:  trace / sum / sum of squares of the covariance (read the way
    libekfcuda reports it: upper triangle mirrored), sum / sum of squares of y, and `nblk` bs x bs blocks of P at fixed
    places (corners, the first landmark's diagonal block, seeded random positions)
    y at the last step, max diagonal entry (the scale of the 1e-9 relative bar), the smallest gate margin seen

configs[1]: 1 000 landmarks x 10 000 steps; configs[2]: 10 000 landmarks x 100 steps; configs[4]: 40 000 landmarks x
100 steps (the oracle's full 80 003^2 covariance is 51 GB: this runs for about an hour on 8 cores and is the reason
the result is a committed fixture instead of a test-time computation).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import StructuredOracle  # noqa: E402
from slam_ros_b200 import scenario as sc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    # capacity = bench.py's (N + headroom): first-fit association now and then matches an aliased landmark (the
    # reference `break`s at the first landmark inside the gate, Robot.cpp:641), the pose jumps, the scan's other lines
    # are appended as new landmarks -- all of which the GPU path has to reproduce, so the capacity (reset threshold,
    # Robot.cpp:893) must be the same on both sides
    "3300": dict(N=3300, steps=25, every=5, nblk=8, bs=32, cap=3300 + 64),
    "1k": dict(N=1000, steps=10000, every=50, nblk=6, bs=16, cap=1000 + 512),
    "10k": dict(N=10000, steps=100, every=10, nblk=12, bs=32, cap=10000 + 1024),
    "40k": dict(N=40000, steps=100, every=10, nblk=12, bs=32, cap=40000 + 1024),
    # north_star's "over 1k steps" at configs[2]'s size, and configs[2]'s batched scans (m = 32, and m = 64 which the GPU path
    # runs as four chunks of 16 lines, each chunk's sweep under the next chunk's line loop)
    "10k_1000": dict(N=10000, steps=1000, every=50, nblk=8, bs=24, cap=10000 + 1024),
    "10k_m32": dict(N=10000, steps=60, every=10, nblk=8, bs=24, cap=10000 + 1024, m=32),
    "10k_m64": dict(N=10000, steps=40, every=10, nblk=8, bs=24, cap=10000 + 1024, m=64),
}
SEED = 1


def block_positions(nl, nblk, bs, seed):
    rng = np.random.default_rng(seed)
    pos = [(0, 0), (0, nl - bs), (nl - bs, nl - bs), (3, 3)]
    while len(pos) < nblk:
        r0, c0 = sorted(int(v) for v in rng.integers(0, nl - bs, 2))
        pos.append((r0, c0))
    return np.array(pos[:nblk], dtype=np.int64)


def checkpoint(so, pos, bs):
    nl = 3 + 2 * so.lines
    Pv = so.P_view()
    tr, sm, sq = so.upper_stats()
    y = np.ctypeslib.as_array(so._lib.ekfo_y_ptr(so._h), shape=(so.n,))[:nl]
    blocks = np.zeros((len(pos), bs, bs))           # entries outside the live part read as zero (ekf_download_block)
    for i, (r0, c0) in enumerate(pos):
        nr, nc = max(0, min(bs, nl - r0)), max(0, min(bs, nl - c0))
        blocks[i, :nr, :nc] = Pv[r0:r0 + nr, c0:c0 + nc]
    return (tr, sm, sq, float(y.sum()), float((y * y).sum())), blocks


def make(name):
    c = CONFIGS[name]
    N, steps, every, bs = c["N"], c["steps"], c["every"], c["bs"]
    M = c.get("m", 8)
    t0 = time.time()
    scn = sc.map_scenario(N, steps, m=M, seed=SEED)
    so = StructuredOracle(c["cap"], threads=0 if N > 2000 else 4)
    so._lib.ekfo_set_threads(so._h, 0 if N > 2000 else 4)
    st, _ = so.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    assert st == 0 and so.lines == N, (st, so.lines)
    nl = 3 + 2 * N
    pos = block_positions(nl, c["nblk"], bs, SEED)
    print("%s: seeded %d landmarks in %.0f s" % (name, N, time.time() - t0), flush=True)
    j_out = np.zeros((steps, M), dtype=np.int32)
    pose = np.zeros((steps, 3))
    lines = np.zeros(steps, dtype=np.int32)
    ck_step, ck_stats, ck_blocks = [], [], []
    for s in range(steps):
        st, j = so.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        assert st == 0, (s, st)
        j_out[s] = j
        pose[s] = so.pose
        lines[s] = so.lines
        if (s + 1) % every == 0 or s == steps - 1:
            stats, blocks = checkpoint(so, pos, bs)
            ck_step.append(s); ck_stats.append(stats); ck_blocks.append(blocks)
            print("%s: step %d  trace %.12g  %.0f s" % (name, s, stats[0], time.time() - t0), flush=True)
    Pv = so.P_view()
    nl = 3 + 2 * so.lines
    diag_max = float(max(Pv[r, r] for r in range(nl)))
    y_last = np.ctypeslib.as_array(so._lib.ekfo_y_ptr(so._h), shape=(so.n,))[:nl].copy()
    info = so.stats()
    out = os.path.join(HERE, "oracle_%s.npz" % name)
    np.savez_compressed(out, N=N, cap=c["cap"], steps=steps, m=M, seed=SEED, every=every, bs=bs, pos=pos, j_out=j_out, pose=pose, lines=lines,
                        ck_step=np.array(ck_step), ck_stats=np.array(ck_stats), ck_blocks=np.array(ck_blocks),
                        y_last=y_last, diag_max=diag_max, min_margin=info["min_margin"], matches=info["matches"])
    print("%s: wrote %s (%.1f KB), min gate margin %.3e, %d matches, %.0f s" %
          (name, out, os.path.getsize(out) / 1e3, info["min_margin"], info["matches"], time.time() - t0), flush=True)
    print("%s: %d landmarks at the end, %d resets" % (name, so.lines, info["resets"]), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for name in (["3300", "10k", "1k", "40k"] if which == "all" else which.split(",")):
        make(name)
