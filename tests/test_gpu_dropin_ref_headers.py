"""The drop-in compiled against the REFERENCE'S OWN headers, driven beside the literal reference by the same harness.

oracle/_ref/libslamdropin.so = dropin/Robot_cuda.cpp (in place of slam_ros/Robot.cpp) + the reference's unmodified
Robot.h / simplifyPath.h / lineFitting.h / lineFitting.cpp / simplifyPath.cpp / vec2.cpp + oracle/ref_harness.cpp,
linked to libekfcuda.so (recipe: oracle/Makefile; built where /root/reference exists, travels to the GPU box).
oracle/_ref/libslamref.so = the same harness over the reference's Robot.cpp.  Both are fed the 1000-step room scenario
with line end points set (line::lineInterval): pose, savedLineCount and Robot::lineIntervals (Robot.cpp:868-879, the
`lines` topic of the node) after every step, getEllipse along the way, y and P_t0 at the end."""
import os

import numpy as np
import pytest

from slam_ros_b200 import scenario as sc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "libslamdropin.so")


def _end_points(z, rng):
    """Two plausible end points (alfa, r) per line in the robot frame, as SetEndPoints would leave them."""
    m = z.shape[0]
    iv = np.zeros((m, 4))
    for i in range(m):
        a, r = z[i]
        for k, off in enumerate((-0.35, 0.4)):
            ang = a + off + 0.05 * rng.standard_normal()
            iv[i, 2 * k] = ang
            iv[i, 2 * k + 1] = r / max(np.cos(off), 0.3)
    return iv


def test_dropin_built_on_reference_headers_matches_the_literal_reference(libekf):
    from oracle.oracle import LiteralReference, have_literal
    if not (have_literal() and os.path.exists(DROPIN)):
        pytest.skip("oracle/_ref/libslamref.so / libslamdropin.so not built (needs /root/reference: make -C oracle ref)")
    steps = 1000
    room = sc.room_scenario(steps=steps, seed=7, range_sigma=5e-5)
    rng = np.random.default_rng(11)
    lit = LiteralReference()
    drp = LiteralReference(so_path=DROPIN)
    assert drp.capacity == lit.capacity == 100 and drp.gate == lit.gate
    n_iv = 0
    worst_pose = worst_iv = 0.0
    resets = 0
    L_prev = 0
    for s in range(steps):
        m = room["count"][s]
        z, R = room["z"][s, :m], room["R"][s, :m]
        iv = _end_points(z, rng)
        _, _, _, pose_l = lit.state(want_cov=False)
        _, _, _, pose_d = drp.state(want_cov=False)
        # the same odometry u for both: each robot gets the encoder pose that makes Robot.cpp:140-145 recover u from ITS pose
        out_l = lit.localize_intervals(z, R, sc.encoder_for(pose_l, room["u"][s]), iv)
        out_d = drp.localize_intervals(z, R, sc.encoder_for(pose_d, room["u"][s]), iv)
        _, _, L_l, pose_l = lit.state(want_cov=False)
        _, _, L_d, pose_d = drp.state(want_cov=False)
        assert L_l == L_d, "savedLineCount at step %d: %d vs %d" % (s, L_l, L_d)
        resets += int(L_l < L_prev); L_prev = L_l
        worst_pose = max(worst_pose, float(np.abs(pose_l - pose_d).max()))
        assert worst_pose < 1e-9, "pose at step %d" % s
        assert out_l.shape == out_d.shape, "lineIntervals length at step %d" % s
        if out_l.size:
            n_iv += out_l.size
            worst_iv = max(worst_iv, float(np.abs(out_l - out_d).max() / max(1.0, np.abs(out_l).max())))
            assert worst_iv < 2e-6, "lineIntervals values at step %d: %s vs %s" % (s, out_l[:8], out_d[:8])   # float32 outputs
        if s % 100 == 99:
            ok_l, ax_l, ang_l = lit.get_ellipse()
            ok_d, ax_d, ang_d = drp.get_ellipse()
            assert ok_l == ok_d and np.allclose(ax_l, ax_d, rtol=1e-5) and abs(ang_l - ang_d) < 1e-4
    y_l, P_l, L_l, _ = lit.state()
    y_d, P_d, L_d, _ = drp.state()                      # the drop-in fills y / P_t0 on demand (Robot::measure)
    nl = 3 + 2 * L_l
    P_lu = np.triu(P_l[:nl, :nl]); P_lu = P_lu + np.triu(P_lu, 1).T      # libekfcuda keeps the upper triangle
    assert np.abs(y_l - y_d).max() / np.abs(y_l).max() < 1e-9
    assert np.abs(P_lu - P_d[:nl, :nl]).max() / np.abs(P_lu).max() < 1e-9
    assert n_iv > 0 and resets > 0
    print("drop-in on the reference's headers vs the literal reference, %d steps: pose %.2e, %d lineIntervals floats (worst %.2e), "
          "%d map resets" % (steps, worst_pose, n_iv, worst_iv, resets))
