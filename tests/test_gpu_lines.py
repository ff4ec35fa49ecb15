"""Line extraction on the device (ekf_lx_*) against the CPU restatement (oracle/lines_oracle.cpp,
itself bitwise equal to the reference's own lineFitting.cpp) and the golden vectors the reference produced.

Tolerances (floating point; the kernel evaluates the fit's pair sums in closed form and reduces in parallel):
  number of lines and their order: exact;  (alfa, r): 1e-10 absolute;  lineInterval end points: 1e-8;
  C_AR: 2e-6 relative -- the reference's own forward differences (eps = 1e-6) carry ~1e-7 relative rounding noise; the
  worst of 11 578 entries over the golden payloads and 240 room scans is 2.2e-7 (scripts/lx_tolerance_probe.py); 2e-5 for
  the deliberately messy payloads (3 cm of range noise, outliers: worst 5.4e-6) and the dense scanners (worst 4.8e-6)."""
import os

import numpy as np
import pytest

from slam_ros_b200 import scenario as sc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lines_literal.npz")


def _compare(rows, n, ref, m, what, car_tol=2e-6):
    assert n == m, "%s: %d lines vs %d" % (what, n, m)
    if n == 0:
        return
    d = np.abs(rows[:, 0] - ref[:, 0])
    d = np.minimum(d, np.abs(d - 2 * np.pi))                      # alfa = +-pi is one direction
    assert d.max() < 1e-10, (what, d.max())
    assert np.abs(rows[:, 1] - ref[:, 1]).max() < 1e-10, what
    assert (rows[:, 3] == 0).all() and (rows[:, 4] == 0).all()
    for c in (2, 5):
        rel = np.abs(rows[:, c] - ref[:, c]) / np.maximum(np.abs(ref[:, c]), 1e-300)
        assert rel.max() < car_tol, (what, c, rel.max())
    # end points: a degenerate segment (duplicated beams: first == last point) gives NaN in the reference too
    assert np.array_equal(np.isnan(rows[:, 6:]), np.isnan(ref[:, 6:])), what
    da = np.abs(rows[:, [6, 8]] - ref[:, [6, 8]])
    da = np.minimum(da, np.abs(da - 2 * 3.14159265))
    assert np.nanmax(da, initial=0.0) < 1e-8, (what, np.nanmax(da, initial=0.0))
    assert np.nanmax(np.abs(rows[:, [7, 9]] - ref[:, [7, 9]]), initial=0.0) < 1e-8, what


def test_golden_payloads_from_the_reference(libekf):
    from slam_ros_b200 import LineExtractor
    g = np.load(GOLD)
    lx = LineExtractor(max_lines=64)
    for s in range(g["scans"].shape[0]):
        rows, n = lx.extract(g["scans"][s])
        m = int(g["count"][s])
        _compare(rows, n, g["rows"][s, :m], m, "golden scan %d" % s)


def test_room_trajectory_against_the_oracle(libekf):
    from slam_ros_b200 import LineExtractor
    from oracle.oracle import LinesOracle
    lo = LinesOracle(); lx = LineExtractor()
    S = sc.room_scans(steps=40, seed=17, range_sigma=2e-3)
    total = 0
    for s in range(40):
        rows, n = lx.extract(S["scans"][s]); ref, m = lo.extract(S["scans"][s])
        _compare(rows, n, ref, m, "scan %d" % s)
        total += n
    assert total > 40 * 15


def test_edge_payloads(libekf):
    """Empty payload, fewer than two returns, tiny segments, dropped beams, and an UNSORTED payload (the reference
    sorts by angle first, lineFitting.cpp:642)."""
    from slam_ros_b200 import LineExtractor
    from oracle.oracle import LinesOracle
    lo = LinesOracle(); lx = LineExtractor()
    full = sc.room_scans(steps=1, seed=5)["scans"][0]
    dead = full.copy(); dead[::2, 0] = 0.0
    rng = np.random.default_rng(3)
    shuffled = full[rng.permutation(full.shape[0])]
    cases = [full[:0], full[:1], full[:3], full[40:47], full[::7], np.concatenate([full[:90], full[200:260]]), dead, shuffled]
    for k, c in enumerate(cases):
        rows, n = lx.extract(c); ref, m = lo.extract(c)
        _compare(rows, n, ref, m, "case %d" % k)


def test_extracted_lines_drive_the_filter(libekf):
    """The node's loop (slam_ros/main.cpp:139-147): payload -> lines -> Robot::localize, all on the device path."""
    from slam_ros_b200 import LineExtractor, EkfFilter
    lx = LineExtractor(); f = EkfFilter(capacity_lines=100)
    S = sc.room_scans(steps=12, seed=9, range_sigma=1e-3)
    for s in range(12):
        rows, n = lx.extract(S["scans"][s])
        m = min(n, 9)                                              # the reference cannot take more than 9 new lines per scan (Q4)
        rc, j, pose = f.scan(S["u"][s], rows[:m, 0:2], rows[:m, 2:6])
        assert rc == 0
    assert f.lines >= 9 and np.isfinite(f.pose).all()


def test_messy_payloads(libekf):
    """Outliers, duplicated beams (equal sort keys), dropped sectors, random order, heavy noise, sparse scans: the
    split decisions and the order of the lines still agree with the oracle on every payload."""
    from slam_ros_b200 import LineExtractor
    from oracle.oracle import LinesOracle
    lx = LineExtractor(); lo = LinesOracle()
    rng = np.random.default_rng(123)
    S = sc.room_scans(steps=30, seed=99, range_sigma=3e-3)["scans"]
    for t in range(90):
        p = S[t % 30].copy()
        kind = t % 6
        if kind == 0:
            k = rng.random(p.shape[0]) < 0.05; p[k, 0] = rng.uniform(0.1, 9.0, k.sum())
        elif kind == 1:
            p = np.concatenate([p, p[rng.integers(0, p.shape[0], 30)]])
        elif kind == 2:
            a = rng.integers(0, 300); p[a:a + rng.integers(5, 60), 0] = 0.0
        elif kind == 3:
            p = p[rng.permutation(p.shape[0])]
        elif kind == 4:
            p[:, 0] += (rng.standard_normal(p.shape[0]) * 0.03).astype(np.float32) * (p[:, 0] > 0)
        else:
            p = p[::rng.integers(2, 6)]
        rows, n = lx.extract(p); ref, m = lo.extract(p)
        _compare(rows, n, ref, m, "payload %d (kind %d)" % (t, kind), car_tol=2e-5)   # 3 cm of range noise: worst 5.4e-6


@pytest.mark.parametrize("beams,step_deg", [(1440, 0.25), (2880, 0.125), (4096, 360.0 / 4096)])
def test_dense_scanners(libekf, beams, step_deg):
    """Up to 4096 returns per payload (the reference's simulator sends 361): segments of several hundred points,
    leaves whose finite-difference covariance the reference computes with O(p^3) trigonometric calls."""
    from slam_ros_b200 import LineExtractor
    from oracle.oracle import LinesOracle
    lo = LinesOracle(); lx = LineExtractor()
    rng = np.random.default_rng(beams)
    for pose in ((0.3, -0.2, 0.1), (-1.0, 0.8, 2.0)):
        p = sc.room_scan(pose, beams=beams, range_sigma=2e-3, rng=rng, step_deg=step_deg)
        rows, n = lx.extract(p); ref, m = lo.extract(p)
        _compare(rows, n, ref, m, "%d beams at %s" % (beams, pose), car_tol=2e-5)   # hundreds of points per leaf: worst 4.8e-6
        assert n >= 15


def test_payload_larger_than_the_limit_is_rejected(libekf):
    from slam_ros_b200 import LineExtractor
    from slam_ros_b200.ekf import EkfError, EKF_EINVAL
    with pytest.raises(EkfError) as e:
        LineExtractor().extract(np.ones((4097, 2), dtype=np.float32))
    assert e.value.code == EKF_EINVAL


def test_world_segments_against_the_reference_lineprovider(libekf):
    """the consumers after the filter.  ekf_lx_world_segments against the reference's own Transform()
    (lineprovider/main.cpp:60-84, compiled unmodified: oracle/_ref/libslamlineprov.so) on the lines the device extracted,
    for poses in every quadrant and across the wrap of theta + PI/2 at pi; scale = 100 is the planner's centimetres
    (astar/main.cpp:44-73: the floats of the message times 100, a float product, so it must be exactly 100 x the scale-1
    result rounded to float)."""
    from slam_ros_b200 import LineExtractor
    from oracle.oracle import LineProviderTransform
    if not LineProviderTransform.available():
        pytest.skip("oracle/_ref/libslamlineprov.so not built (needs /root/reference: make -C oracle ref)")
    ref = LineProviderTransform()
    lx = LineExtractor()
    S = sc.room_scans(steps=12, seed=41, range_sigma=2e-3)
    poses = [(0.0, 0.0, 0.0), (1.5, -2.25, 0.4), (-3.0, 0.7, 1.6), (0.2, 0.1, 1.5707963), (2.0, 2.0, 3.1), (-1.0, -1.0, -2.9),
             (4.0, -0.5, -1.0), (0.3, 0.3, 2.0), (0.0, 5.0, 1.58), (7.0, 1.0, -0.01), (-2.0, 3.0, 3.14159), (1.0, 1.0, -3.14)]
    total = 0
    for s, pose in enumerate(poses):
        rows, n = lx.extract(S["scans"][s])
        seg = lx.world_segments(pose)
        assert seg.shape == (n, 4) and n > 5
        ok = ~np.isnan(rows[:, 6:10]).any(axis=1)                # degenerate segments carry NaN end points in the reference too
        want = ref.transform(rows[ok, 6:10], pose)
        assert np.abs(seg[ok] - want).max() <= 2e-6 * max(1.0, float(np.abs(want).max())), (s, np.abs(seg[ok] - want).max())
        assert np.isnan(seg[~ok]).all()
        cm = lx.world_segments(pose, scale=100.0)
        assert np.array_equal(cm[ok], seg[ok] * np.float32(100.0))
        total += int(ok.sum())
    assert total > 100
