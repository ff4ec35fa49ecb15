"""Host-side mirror of the slam_ros node's callbacks and loop body (slam_ros/main.cpp:37-177) without ROS.

The reference node is three subscriber callbacks that fill globals and a 10 Hz loop that, when a scan and an
external pose are pending, runs `rover->localize(lines, NULL, encoderPose)` and publishes two messages.  This
class keeps that state machine -- same names, same order of effects -- over the device path: `mapping_cb` runs
the line extraction (ekf_lx_*), `spin_once` the EKF (ekf_scan) and returns what the node would publish:

    robotPosition : geometry_msgs/Transform  translation = (x, y, theta); rotation.x/.y/.z = major, minor, angle
    lines         : std_msgs/Float32MultiArray  4 floats (x0, y0, x1, y1) per line appended to the map

It is plumbing for end-to-end parity tests and for SURVEY 8f row 3; a ROS wrapper only has to forward messages.
"""
import numpy as np

from .ekf import LineExtractor
from .robot import Robot, Line, LINESIZE


class SlamNode:
    def __init__(self, device=0, max_new_lines=None, rover=None, extractor=None):
        # rover / extractor: injected stand-ins for host-logic tests; by default the device objects
        self.rover = rover if rover is not None else Robot(0.0, 0.0, 0.0, device=device)          # main.cpp:98
        self.extractor = extractor if extractor is not None else LineExtractor(device=device)
        self.lines = []                                            # main.cpp:35
        self.encoderPose = [0.0, 0.0, 0.0]                         # main.cpp:32
        self.rot = [0.0, 0.0]
        self.update = False
        self.encoderUpdate = False
        self.sensUpdate = False
        # The reference appends every unmatched line and only then tests for the map reset, overrunning y[] when a
        # scan brings more lines than the head-room; libekfcuda drops what does not fit and reports
        # EKF_ECAPACITY.  max_new_lines lets a harness keep BOTH sides inside the defined behaviour.
        self.max_new_lines = max_new_lines
        self.last_status = 0

    # ---- subscribers ---------------------------------------------------------------------------
    def mapping_cb(self, data):
        """main.cpp:37-78 (SIMULATIONOFF branch): payload of (r, angle) float pairs -> self.lines."""
        rows, n = self.extractor.extract(np.asarray(data, dtype=np.float32))
        self.lines = [Line(alfa=float(r[0]), r=float(r[1]), C_AR=tuple(float(v) for v in r[2:6]),
                           lineInterval=[(float(r[6]), float(r[7])), (float(r[8]), float(r[9]))]) for r in rows]
        self.sensUpdate = True

    def encoderUpdate_cb(self, x, y):
        """main.cpp:79-83"""
        self.rot = [float(x), float(y)]
        self.update = True

    def realpose_cb(self, x, y, z):
        """main.cpp:84-89"""
        self.encoderPose = [float(x), float(y), float(z)]
        self.encoderUpdate = True

    # ---- one pass of the while(ros::ok()) body, main.cpp:130-176 ----------------------------------
    def spin_once(self):
        """Returns None when nothing was pending, else (robotPosition dict, lines float32 array)."""
        if self.sensUpdate:                                        # `if(encoderPose && sensUpdate)`: the array is never null
            self.update = True
        if not self.update:
            return None
        self.update = False
        self.encoderUpdate = False
        self.sensUpdate = False
        lines = self.lines if self.max_new_lines is None else self.lines[:self.max_new_lines]
        self.last_status = self.rover.localize(lines, None, self.encoderPose)       # main.cpp:144
        self.lines = []
        msg = self.rover.robotPosition()                           # main.cpp:150-169
        out_lines = np.asarray(self.rover.lineIntervals, dtype=np.float32)          # main.cpp:172
        self.rover.lineIntervals = []                              # main.cpp:174
        return msg, out_lines


__all__ = ["SlamNode", "LINESIZE"]
