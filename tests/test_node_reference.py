"""the reference NODE itself.  slam_ros/main.cpp is compiled UNMODIFIED (oracle/Makefile, target
_ref/libslamnode_ref.so) over an in-process restatement of the roscpp calls it makes (oracle/ros_shim/), with the
reference's own Robot.cpp, lineFitting.cpp, simplifyPath.cpp, vec2.cpp.  Here (CPU) the node's published messages are checked
against the pipeline made of the same pieces driven directly -- the reference's LineExtraction feeding the structured
oracle of Robot::localize -- which pins the harness; tests/test_gpu_node_dropin.py then runs the same node with the
drop-in Robot over libekfcuda.so."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from node_driver import NODE_REF, run_node  # noqa: E402
from slam_ros_b200 import scenario as sc  # noqa: E402


def _run_in_child(code):
    """main.cpp keeps its state in globals: one run per process."""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_reference_node_publishes_what_its_pieces_compute(tmp_path):
    if not os.path.exists(NODE_REF):
        pytest.skip("oracle/_ref/libslamnode_ref.so not built (needs /root/reference: make -C oracle ref)")
    from oracle.oracle import LiteralLineExtraction, StructuredOracle, have_literal_lines
    if not have_literal_lines():
        pytest.skip("oracle/_ref/libslamlines.so not built")
    steps = 60
    out = str(tmp_path / "node.npz")
    _run_in_child(
        "import sys; sys.path.insert(0, 'tests'); import numpy as np\n"
        "from node_driver import NODE_REF, run_node\n"
        "from slam_ros_b200 import scenario as sc\n"
        "S = sc.room_scans(steps=%d, seed=31, range_sigma=1e-3, d=0.03)\n"
        "S['scans'][:, 110:, 0] = 0.0\n"
        "poses, lines = run_node(NODE_REF, S['scans'], S['u'], sc.encoder_for)\n"
        "np.savez(%r, poses=poses, counts=np.array([l.size for l in lines]), lines=np.concatenate(lines) if lines else np.zeros(0))\n"
        % (steps, out))
    got = np.load(out)
    assert got["poses"].shape == (steps, 6)
    S = sc.room_scans(steps=steps, seed=31, range_sigma=1e-3, d=0.03)
    # Only the first 110 beams return: at most 9 lines per scan, so that the map (reset beyond 90 lines, Robot.cpp:893) never
    # outgrows y[203] -- the reference node has no guard there and would write past the array.
    S["scans"][:, 110:, 0] = 0.0
    ex = LiteralLineExtraction()
    so = StructuredOracle(100)
    total_new = 0
    for s in range(steps):
        rows, n = ex.extract(S["scans"][s])
        L0 = so.lines
        so.localize(rows[:, 0:2], rows[:, 2:6], sc.encoder_for(so.pose, S["u"][s]))
        assert np.abs(got["poses"][s, :3] - so.pose).max() < 1e-9, "pose at step %d" % s
        new = max(so.lines - L0, 0)
        assert got["counts"][s] == 4 * new or so.lines < L0, "lines message at step %d" % s     # (a map reset empties the map)
        total_new += new
    ok, ax, ang = so.get_ellipse()
    assert ok and np.allclose(got["poses"][-1, 3:5], (ax[1], ax[0]), rtol=1e-5) and abs(got["poses"][-1, 5] - ang) < 1e-5
    assert total_new >= 100 and so.stats()["resets"] >= 3 and so.stats()["matches"] >= 80
