"""Generates tests/golden/literal_1k.npz by running the REFERENCE ITSELF at BASELINE configs[1] size:
oracle/_ref/libslamref1k.so = slam_ros/Robot.cpp compiled with LINESIZE = 1000 / SLAMSIZE = 2003 (the two macros of
Robot.h:13-14 rewritten into a generated header by oracle/Makefile; Q1-patched on a pipe; GSL shim).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_1k.py   (~40 s: every
Robot::localize call is a dense 2003^3 dgemm).  Inputs are stored next to the outputs; of the 2003 x 2003 covariance
the fixture keeps the diagonal, the row sums, the robot rows and 12 blocks of 48 x 48 spread over the matrix."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import LiteralReference, build  # noqa: E402
from slam_ros_b200 import scenario as sc  # noqa: E402

N, STEPS, M = 600, 2, 8
BLOCK = 48


def corners(n_live):
    rng = np.random.default_rng(0)
    c = [(0, 0), (0, n_live - BLOCK), (n_live - BLOCK, n_live - BLOCK), (3, 3)]
    c += [tuple(int(v) for v in rng.integers(0, n_live - BLOCK, 2)) for _ in range(8)]
    return np.array(c, dtype=np.int64)


def main():
    build()
    lit = LiteralReference(big=True)
    scn = sc.map_scenario(N, STEPS, m=M, seed=41)
    zero = np.zeros(3)
    lit.localize(scn["seed_z"], scn["seed_R"], sc.encoder_for(zero, zero))
    poses = np.zeros((STEPS, 3)); enc = np.zeros((STEPS, 3))
    for s in range(STEPS):
        y, P, L, pose = lit.state()
        enc[s] = sc.encoder_for(pose, scn["u"][s])
        lit.localize(scn["z"][s], scn["R"][s], enc[s])
        poses[s] = lit.state()[3]
    y, P, L, pose = lit.state()
    nl = 3 + 2 * L
    cs = corners(nl)
    blocks = np.stack([P[r:r + BLOCK, c:c + BLOCK] for r, c in cs])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "literal_1k.npz")
    np.savez_compressed(out, seed_z=scn["seed_z"], seed_R=scn["seed_R"], u=scn["u"], z=scn["z"], R=scn["R"], encoder=enc,
                        pose=poses, L=np.array([L]), y=y[:nl], diag=np.diag(P)[:nl].copy(), rowsum=P[:nl, :nl].sum(axis=1),
                        top=P[:3, :nl].copy(), corners=cs, blocks=blocks, pmax=np.array([np.abs(P).max()]))
    print("wrote", out, "L", L, "pose", pose)


if __name__ == "__main__":
    main()
