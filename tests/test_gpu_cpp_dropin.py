"""The C++ host side (include/ekf_robot.hpp) compiled with g++ against libekfcuda.so and driven like the
reference node drives Robot::localize; results checked against the CPU oracle."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from slam_ros_b200 import scenario as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_robot_dropin_matches_oracle(libekf):
    from oracle.oracle import StructuredOracle
    from slam_ros_b200 import library_path
    steps = 150
    room = sc.room_scenario(steps=steps, range_sigma=5e-5)
    maxl = room["z"].shape[1]
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "dropin"); fin = os.path.join(d, "in.bin"); fout = os.path.join(d, "out.bin")
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), library_path(),
                               "-Wl,-rpath," + os.path.dirname(library_path()), "-o", exe])
        with open(fin, "wb") as f:
            np.array([steps, maxl], dtype=np.int32).tofile(f)
            room["count"].astype(np.int32).tofile(f)
            room["u"].astype(np.float64).tofile(f)
            room["z"].astype(np.float64).tofile(f)
            room["R"].astype(np.float64).tofile(f)
        subprocess.check_call([exe, fin, fout])
        out = np.fromfile(fout, dtype=np.float64)
    rec = out[:4 * steps].reshape(steps, 4)
    ell = out[4 * steps:4 * steps + 4]
    n = int(out[4 * steps + 4])
    y = out[4 * steps + 5:4 * steps + 5 + n]
    P = out[4 * steps + 5 + n:4 * steps + 5 + n + n * n].reshape(n, n)
    n_intervals = int(out[-1])
    so = StructuredOracle(100)
    added = 0
    for s in range(steps):
        m = room["count"][s]
        L0 = so.lines
        enc = sc.encoder_for(so.pose, room["u"][s])
        st, j = so.localize(room["z"][s, :m], room["R"][s, :m], enc)
        assert so.lines == int(rec[s, 3]), "line count at step %d" % s
        assert np.abs(so.pose - rec[s, :3]).max() < 1e-9, "pose at step %d" % s
        added += int((j < 0).sum())
    assert n == so.n
    assert np.abs(y - so.y_full()).max() / np.abs(so.y_full()).max() < 1e-9
    assert np.abs(P - so.P_full()).max() / np.abs(so.P_full()).max() < 1e-9
    ok, ax, ang = so.get_ellipse()
    assert ell[0] == 1.0 and np.allclose(ell[1:3], ax, rtol=1e-5)
    assert n_intervals == 4 * added          # two end points (x, y) per appended line, Robot.cpp:869-879


def test_cpp_line_extractor_feeds_robot(libekf):
    """ekfcuda::LineExtractor + ekfcuda::Robot compiled with g++ and driven like the node's callback and loop
    (slam_ros/main.cpp:37-71, 139-174); the extracted lines are checked against the CPU oracle."""
    from oracle.oracle import LinesOracle
    from slam_ros_b200 import library_path
    steps = 6
    S = sc.room_scans(steps=steps, seed=21, range_sigma=1e-3)
    beams = S["scans"].shape[1]
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "lines"); fin = os.path.join(d, "in.bin"); fout = os.path.join(d, "out.bin")
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", "lines_main.cpp"), library_path(),
                               "-Wl,-rpath," + os.path.dirname(library_path()), "-o", exe])
        with open(fin, "wb") as f:
            np.array([steps, beams], dtype=np.int32).tofile(f)
            S["scans"].astype(np.float32).tofile(f)
            S["u"].astype(np.float64).tofile(f)
        subprocess.check_call([exe, fin, fout])
        out = np.fromfile(fout, dtype=np.float64)
    lo = LinesOracle()
    k = 0
    for s in range(steps):
        n = int(out[k]); k += 1
        rows = out[k:k + 10 * n].reshape(n, 10); k += 10 * n
        pose = out[k:k + 4]; k += 4
        ref, m = lo.extract(S["scans"][s])
        assert n == m
        assert np.abs(rows[:, :2] - ref[:, :2]).max() < 1e-10
        assert (np.abs(rows[:, [2, 5]] - ref[:, [2, 5]]) / ref[:, [2, 5]]).max() < 2e-6
        assert np.abs(rows[:, 6:] - ref[:, 6:]).max() < 1e-8
        assert np.isfinite(pose).all() and pose[3] >= 9
        seg = out[k:k + 4 * n].reshape(n, 4); k += 4 * n          # LineExtractor::worldSegments at the pose just published
        ok = ~np.isnan(rows[:, 6:10]).any(axis=1)
        th = pose[2]
        for e in (0, 1):                                          # lineprovider/main.cpp:60-84 in numpy (float output: 1e-5)
            px = rows[ok, 7 + 2 * e] * np.cos(rows[ok, 6 + 2 * e]); py = rows[ok, 7 + 2 * e] * np.sin(rows[ok, 6 + 2 * e])
            wx = np.cos(th) * px - np.sin(th) * py + pose[0]; wy = np.sin(th) * px + np.cos(th) * py + pose[1]
            assert np.abs(seg[ok, 2 * e] - wx).max() < 1e-5 and np.abs(seg[ok, 2 * e + 1] - wy).max() < 1e-5
    assert k == out.size
