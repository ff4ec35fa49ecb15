"""End to end through the node mirror (slam_ros_b200.SlamNode = slam_ros/main.cpp:37-177 without ROS):
`mappingPoints` payload -> device line extraction -> device EKF -> `robotPosition` / `lines` messages,
against the CPU pipeline made of the reference's own pieces: its LineExtraction (oracle/_ref/libslamlines.so where
built, else the bitwise-equal restatement) feeding the structured oracle of Robot::localize (bitwise equal to the
reference's Robot.cpp).  Tolerance: the two extractions differ by ~1e-12 in (alfa, r) and ~1e-7 relative in C_AR,
which the filter turns into pose differences far below 1e-6; the map size must agree exactly."""
import numpy as np
import pytest

from slam_ros_b200 import scenario as sc

pytestmark = pytest.mark.gpu


def test_payload_to_pose_matches_the_reference_pipeline(libekf):
    from slam_ros_b200 import SlamNode
    from oracle.oracle import LinesOracle, LiteralLineExtraction, StructuredOracle, have_literal_lines
    steps = 60
    S = sc.room_scans(steps=steps, seed=31, range_sigma=1e-3, d=0.03)
    node = SlamNode(max_new_lines=9)                    # both sides stay inside the reference's defined behaviour (Q4)
    ex = LiteralLineExtraction() if have_literal_lines() else LinesOracle()
    so = StructuredOracle(100)
    published = 0
    for s in range(steps):
        u = S["u"][s]
        node.realpose_cb(*sc.encoder_for((node.rover.xPos, node.rover.yPos, node.rover.thetaPos), u))
        node.mapping_cb(S["scans"][s])
        out = node.spin_once()
        assert out is not None and node.spin_once() is None          # nothing pending afterwards
        msg, line_msg = out
        rows, n = ex.extract(S["scans"][s])
        rows = rows[:9]
        so.localize(rows[:, 0:2], rows[:, 2:6], sc.encoder_for(so.pose, u))
        assert node.rover.savedLineCount == so.lines, "map size at step %d" % s
        assert np.abs(np.array(msg["translation"]) - so.pose).max() < 1e-6, "pose at step %d" % s
        assert line_msg.size % 4 == 0
        published += line_msg.size // 4
    ok, ax, ang = so.get_ellipse()
    assert msg["ellipse_ok"] == ok and np.allclose(msg["rotation"][:2], (ax[1], ax[0]), rtol=1e-4)
    assert published >= 20                               # end points of every appended line went out
