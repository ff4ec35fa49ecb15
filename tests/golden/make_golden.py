"""Generates tests/golden/room_literal.npz by running the REFERENCE ITSELF (oracle/_ref/libslamref.so =
slam_ros/Robot.cpp, Q1-patched on a pipe, compiled over oracle/gsl_shim) on the configs[0] room scenario.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The inputs are stored next to the outputs so the fixture does not depend on the scenario generator.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import LiteralReference, build  # noqa: E402
from slam_ros_b200 import scenario as sc  # noqa: E402

STEPS = 300


def main():
    build()
    room = sc.room_scenario(steps=STEPS, seed=7, range_sigma=5e-5)
    lit = LiteralReference()
    poses = np.zeros((STEPS, 3)); Ls = np.zeros(STEPS, dtype=np.int64)
    traces = np.zeros(STEPS); sums = np.zeros(STEPS); encs = np.zeros((STEPS, 3))
    snaps = {}
    for s in range(STEPS):
        m = room["count"][s]
        y, P, L, pose = lit.state()
        enc = sc.encoder_for(pose, room["u"][s])
        encs[s] = enc
        lit.localize(room["z"][s, :m], room["R"][s, :m], enc)
        y, P, L, pose = lit.state()
        poses[s] = pose; Ls[s] = L; traces[s] = np.trace(P); sums[s] = P.sum()
        if s in (0, 1, 10, 99, STEPS - 1):
            snaps["y_%d" % s] = y.copy(); snaps["P_%d" % s] = P.copy()
    ok, ax, ang = lit.get_ellipse()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "room_literal.npz")
    np.savez_compressed(out, u=room["u"], z=room["z"], R=room["R"], count=room["count"], encoder=encs,
                        pose=poses, L=Ls, trace=traces, psum=sums, ellipse=np.array([ax[0], ax[1], ang]), **snaps)
    print("wrote", out, os.path.getsize(out), "bytes; final L", Ls[-1], "max L", Ls.max(), "range errors", lit.range_errors())


if __name__ == "__main__":
    main()
