"""Generates tests/golden/lines_literal.npz by running the REFERENCE ITSELF (oracle/_ref/libslamlines.so =
slam_ros/lineFitting.cpp + simplifyPath.cpp + vec2.cpp compiled where they lie, zero-initialised automatic
variables and gsl_matrix_alloc storage -- see oracle/ref_lines_harness.cpp) on `mappingPoints` payloads of the
synthetic room.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_lines.py
The payloads are stored next to the outputs so the fixture does not depend on the scenario generator.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import LiteralLineExtraction, build  # noqa: E402
from slam_ros_b200 import scenario as sc  # noqa: E402

STEPS = 8
MAX_LINES = 64


def main():
    build()
    lit = LiteralLineExtraction()
    S = sc.room_scans(steps=STEPS, seed=11, range_sigma=1e-3)
    rows = np.zeros((STEPS, MAX_LINES, 10)); counts = np.zeros(STEPS, dtype=np.int64)
    for s in range(STEPS):
        r, n = lit.extract(S["scans"][s], MAX_LINES)
        rows[s, :len(r)] = r; counts[s] = n
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lines_literal.npz")
    np.savez_compressed(out, scans=S["scans"], rows=rows, count=counts)
    print("wrote", out, "lines per scan:", counts.tolist())


if __name__ == "__main__":
    main()
