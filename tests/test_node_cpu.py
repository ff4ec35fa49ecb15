"""Host logic of the node mirror (slam_ros_b200/node.py = slam_ros/main.cpp:37-177) with stand-ins for the two
device objects: the order of effects of the callbacks and of one pass of the loop body."""
import numpy as np

from slam_ros_b200.node import SlamNode


class _FakeExtractor:
    def __init__(self, n):
        self.n = n; self.calls = 0

    def extract(self, payload):
        self.calls += 1
        rows = np.zeros((self.n, 10))
        rows[:, 0] = np.linspace(-1, 1, self.n); rows[:, 1] = 2.0 + np.arange(self.n)
        rows[:, 2] = 1e-5; rows[:, 5] = 2e-5
        rows[:, 6:] = 0.5
        return rows, self.n


class _FakeRover:
    def __init__(self):
        self.xPos = self.yPos = self.thetaPos = 0.0
        self.lineIntervals = []
        self.seen = []

    def localize(self, lines, rot, encoder):
        self.seen.append((len(lines), rot, tuple(encoder)))
        self.xPos += 1.0
        self.lineIntervals += [1.0, 2.0, 3.0, 4.0] * len(lines)
        return 0

    def robotPosition(self):
        return {"translation": (self.xPos, self.yPos, self.thetaPos), "rotation": (0.2, 0.1, 0.3), "ellipse_ok": True}


def test_loop_body_order_of_effects():
    rover = _FakeRover(); ex = _FakeExtractor(12)
    node = SlamNode(rover=rover, extractor=ex, max_new_lines=9)
    assert node.spin_once() is None                              # nothing pending (main.cpp:139-142)
    node.realpose_cb(0.5, -0.5, 0.1)
    assert node.spin_once() is None                              # a pose alone does not trigger an update
    node.mapping_cb(np.zeros((361, 2), dtype=np.float32))
    assert ex.calls == 1 and len(node.lines) == 12 and node.sensUpdate
    assert node.lines[3].C_AR == (1e-5, 0.0, 0.0, 2e-5) and len(node.lines[3].lineInterval) == 2
    msg, out = node.spin_once()
    assert rover.seen == [(9, None, (0.5, -0.5, 0.1))]           # rot = NULL, encoderPose (main.cpp:144); capped at 9
    assert msg["translation"][0] == 1.0 and out.dtype == np.float32 and out.size == 36
    assert node.lines == [] and rover.lineIntervals == []        # main.cpp:147, 174
    assert not (node.update or node.sensUpdate or node.encoderUpdate)
    assert node.spin_once() is None
    node.encoderUpdate_cb(0.1, 0.2)                              # the encoder callback alone forces an update (main.cpp:79-83)
    msg, out = node.spin_once()
    assert rover.seen[-1][0] == 0 and out.size == 0
