"""Line extraction: the CPU restatement oracle/lines_oracle.cpp pinned
(i) BITWISE against the reference's own lineFitting.cpp / simplifyPath.cpp (deterministic build,
    oracle/_ref/libslamlines.so) -- only where /root/reference was available to build it, and
(ii) against golden vectors that build produced (tests/golden/lines_literal.npz, make_golden_lines.py)."""
import os

import numpy as np
import pytest

from oracle.oracle import LinesOracle, LiteralLineExtraction, have_literal_lines
from slam_ros_b200 import scenario as sc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lines_literal.npz")


def test_restatement_matches_golden_vectors_bitwise():
    g = np.load(GOLD)
    lo = LinesOracle()
    for s in range(g["scans"].shape[0]):
        rows, n = lo.extract(g["scans"][s], 64)
        assert n == int(g["count"][s])
        assert np.array_equal(rows, g["rows"][s, :n]), "scan %d" % s
    assert g["count"].min() >= 15           # the room gives ~20 lines per scan


@pytest.mark.skipif(not have_literal_lines(), reason="literal reference not built (no /root/reference)")
def test_restatement_matches_literal_reference_bitwise():
    lo = LinesOracle(); lit = LiteralLineExtraction()
    S = sc.room_scans(steps=10, seed=29, range_sigma=2e-3)
    for s in range(10):
        a, n = lo.extract(S["scans"][s]); b, m = lit.extract(S["scans"][s])
        assert n == m and np.array_equal(a, b), "scan %d" % s


@pytest.mark.skipif(not have_literal_lines(), reason="literal reference not built (no /root/reference)")
def test_edge_payloads_match_literal_reference():
    lo = LinesOracle(); lit = LiteralLineExtraction()
    S = sc.room_scans(steps=1, seed=5)
    full = S["scans"][0]
    cases = [full[:0], full[:1], full[:3], full[40:47], full[::7], np.concatenate([full[:90], full[200:260]])]
    dead = full.copy(); dead[::2, 0] = 0.0                   # every other beam without a return (r <= 0.05 is dropped)
    cases.append(dead)
    for k, c in enumerate(cases):
        a, n = lo.extract(c); b, m = lit.extract(c)
        assert n == m and np.array_equal(a, b), "case %d" % k


def test_lines_are_canonical_and_feed_the_filter_convention():
    """(alfa, r) with r >= 0 and alfa in (-pi, pi], C_AR diagonal with positive variances below the 0.01 cut
    (lineFitting.cpp:586-638, main.cpp:66-69) -- what Robot::localize consumes (simplifyPath.h:62-79)."""
    lo = LinesOracle()
    S = sc.room_scans(steps=3, seed=2)
    for s in range(3):
        rows, n = lo.extract(S["scans"][s])
        assert n == len(rows) and n > 10
        assert (rows[:, 1] >= 0).all() and (np.abs(rows[:, 0]) <= np.pi + 1e-12).all()
        assert (rows[:, 3] == 0).all() and (rows[:, 4] == 0).all()
        assert (rows[:, 2] >= 0).all() and (rows[:, 2] <= 0.01).all() and (rows[:, 5] >= 0).all()


def test_closed_form_of_the_pair_sums_used_on_the_device():
    """The identity k_lx_segments relies on (slam_ros_b200/csrc/ekf_lines.cu): with X = r cos a, Y = r sin a and unit
    weights, (2/p) sum_1 + (1/p) sum_2 = -2 S_xy and (2/p) sum_3 + (1/p) sum_4 = -(S_xx - S_yy) (centred moments),
    and sum_i r_i cos(a_i - alfa) = cos(alfa) SX + sin(alfa) SY -- so the O(p) evaluation returns the fit of
    lineFitting.cpp:267-304 (checked here against the restatement's O(p^2) pair sums)."""
    lo = LinesOracle()
    rng = np.random.default_rng(4)
    PI = 3.14159265
    for _ in range(40):
        p = int(rng.integers(2, 120))
        al = rng.uniform(-np.pi, np.pi); R = rng.uniform(0.5, 8.0)
        th = al + np.sort(rng.uniform(-0.6, 0.6, p))
        r = R / np.cos(th - al) + rng.standard_normal(p) * 3e-3
        ref = lo.fit(th, r)
        X = r * np.cos(th); Y = r * np.sin(th)
        dx = X - X.mean(); dy = Y - Y.mean()
        ar = 0.5 * np.arctan2(-2.0 * (dx * dy).sum(), -((dx * dx).sum() - (dy * dy).sum()))
        rr = (np.cos(ar) * X.sum() + np.sin(ar) * Y.sum()) / p
        assert abs((ar * 180 / PI) * (PI / 180) - ref[0]) < 1e-11 and abs(rr - ref[1]) < 1e-11


def test_lineprovider_transform_harness():
    """The reference's own Transform() (lineprovider/main.cpp:60-84, oracle/_ref/libslamlineprov.so) behaves as the rotation
    plus translation it is meant to be -- including its quirk: ey is built from theta + PI/2 with PI = 3.14159265
    (lineFitting.h:12), so the frame is orthogonal only to ~2e-9 -- which pins the harness that the device path is
    compared with (tests/test_gpu_lines.py)."""
    import numpy as np
    import pytest
    from oracle.oracle import LineProviderTransform
    if not LineProviderTransform.available():
        pytest.skip("oracle/_ref/libslamlineprov.so not built (needs /root/reference: make -C oracle ref)")
    ref = LineProviderTransform()
    rng = np.random.default_rng(3)
    iv = np.column_stack([rng.uniform(-3.1, 3.1, 200), rng.uniform(0.1, 9.0, 200), rng.uniform(-3.1, 3.1, 200), rng.uniform(0.1, 9.0, 200)])
    for pose in ((0.0, 0.0, 0.0), (1.0, -2.0, 0.7), (-4.0, 3.0, 2.9), (0.5, 0.5, -3.0), (2.0, 1.0, 1.5707963)):
        out = ref.transform(iv, pose)
        th = pose[2]
        for e in (0, 1):
            px = iv[:, 1 + 2 * e] * np.cos(iv[:, 2 * e]); py = iv[:, 1 + 2 * e] * np.sin(iv[:, 2 * e])
            wx = np.cos(th) * px - np.sin(th) * py + pose[0]; wy = np.sin(th) * px + np.cos(th) * py + pose[1]
            assert np.abs(out[:, 2 * e] - wx).max() < 2e-6 and np.abs(out[:, 2 * e + 1] - wy).max() < 2e-6
