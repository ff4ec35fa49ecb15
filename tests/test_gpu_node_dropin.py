"""SURVEY 8f row 3, end to end: the reference NODE -- slam_ros/main.cpp compiled UNMODIFIED over oracle/ros_shim -- run twice
on the same closed-loop message script: once with the reference's own Robot.cpp (oracle/_ref/libslamnode_ref.so) and once
with the drop-in, dropin/Robot_cuda.cpp over libekfcuda.so (oracle/_ref/libslamnode_dropin.so).  Everything the node
publishes is compared: every `robotPosition` message (pose = translation, uncertainty ellipse = rotation) and every `lines`
message (four floats per line appended to the map).  60 scans, 4 map resets, ~100 matched and ~390 appended lines."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from node_driver import NODE_DROPIN, NODE_REF  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = """
import sys; sys.path.insert(0, 'tests'); import numpy as np
from node_driver import run_node
from slam_ros_b200 import scenario as sc
S = sc.room_scans(steps=%d, seed=31, range_sigma=1e-3, d=0.03)
S['scans'][:, 110:, 0] = 0.0        # at most 9 lines per scan: the reference node stays inside y[203] (SURVEY Q4)
poses, lines = run_node(%r, S['scans'], S['u'], sc.encoder_for)
np.savez(%r, poses=poses, counts=np.array([l.size for l in lines]), lines=np.concatenate(lines) if lines else np.zeros(0))
"""


def run_child(so_path, steps, out):
    r = subprocess.run([sys.executable, "-c", CHILD % (steps, so_path, out)], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    return np.load(out)


def test_reference_node_with_the_dropin_publishes_the_same_messages(libekf, tmp_path):
    if not (os.path.exists(NODE_REF) and os.path.exists(NODE_DROPIN)):
        pytest.skip("oracle/_ref/libslamnode_{ref,dropin}.so not built (needs /root/reference: make -C oracle ref)")
    steps = 60
    lit = run_child(NODE_REF, steps, str(tmp_path / "lit.npz"))
    dro = run_child(NODE_DROPIN, steps, str(tmp_path / "dropin.npz"))
    assert lit["poses"].shape == dro["poses"].shape == (steps, 6)
    assert np.array_equal(lit["counts"], dro["counts"])                   # same lines appended at every step: same associations
    assert lit["counts"].sum() >= 4 * 300
    assert np.abs(lit["poses"][:, :3] - dro["poses"][:, :3]).max() < 1e-9   # pose after every scan
    # the ellipse goes through float and an eigen-decomposition in the reference (Robot.cpp:73-124)
    assert np.allclose(lit["poses"][:, 3:5], dro["poses"][:, 3:5], rtol=1e-4, atol=1e-7)
    dang = np.abs(lit["poses"][:, 5] - dro["poses"][:, 5])
    assert np.minimum(dang, np.abs(dang - np.pi)).max() < 1e-3
    assert np.abs(lit["lines"] - dro["lines"]).max() < 1e-5              # end points of the appended lines (floats)
