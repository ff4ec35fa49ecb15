"""Drives the reference NODE -- slam_ros/main.cpp compiled unmodified over oracle/ros_shim (oracle/node_harness.cpp) -- from
Python: a closed loop in which every tick supplies `realRoboPose` (built from the node's latest `robotPosition`, as the
simulator's feedback would be) and one `mappingPoints` payload.  Test infrastructure."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE_REF = os.path.join(ROOT, "oracle", "_ref", "libslamnode_ref.so")
NODE_DROPIN = os.path.join(ROOT, "oracle", "_ref", "libslamnode_dropin.so")

TICK_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.POINTER(C.c_float)),
                      C.POINTER(C.c_int))


def run_node(so_path, scans, u, encoder_for):
    """Runs the node library at so_path over the payloads `scans` (S, beams, 2) float32 and odometry u (S, 3).
    Returns (poses (S, 6): translation xyz + rotation xyz of every robotPosition message, [lines message per step])."""
    lib = C.CDLL(so_path, mode=C.RTLD_LOCAL)
    lib.slam_node_run.argtypes = [TICK_FN, C.c_int]
    lib.slam_node_get_pose.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.slam_node_get_lines.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int]
    steps = scans.shape[0]
    payloads = [np.ascontiguousarray(scans[s], dtype=np.float32).reshape(-1) for s in range(steps)]

    def tick(t, last_pose, n_pub, real_pose, points, n_floats):
        if t >= steps:
            return 0
        pose = (last_pose[0], last_pose[1], last_pose[2]) if last_pose else (0.0, 0.0, 0.0)
        enc = encoder_for(pose, u[t])
        real_pose[0], real_pose[1], real_pose[2] = float(enc[0]), float(enc[1]), float(enc[2])
        points[0] = payloads[t].ctypes.data_as(C.POINTER(C.c_float))
        n_floats[0] = payloads[t].size
        return 1

    cb = TICK_FN(tick)
    rc = lib.slam_node_run(cb, steps + 1)
    assert rc == 0
    n = lib.slam_node_published()
    poses = np.zeros((n, 6))
    lines = []
    buf = (C.c_float * 4096)()
    for i in range(n):
        lib.slam_node_get_pose(i, poses[i].ctypes.data_as(C.POINTER(C.c_double)))
        k = lib.slam_node_get_lines(i, buf, 4096)
        lines.append(np.array(buf[:k], dtype=np.float32))
    return poses, lines
