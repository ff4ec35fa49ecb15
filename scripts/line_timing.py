"""Phase timestamps of the line-loop kernel (library built with EKF_NVCC_EXTRA=-DEKF_LINE_TIMING)."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc, load_library
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
scn = sc.map_scenario(N, 6, m=8, seed=1)
f = EkfFilter(capacity_lines=N + 256)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
for s in range(6):
    f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
lib = load_library()
buf = (C.c_ulonglong * (16 * 16))()
print("rc", lib.ekf_debug_line_timing(buf, 16 * 16))
t = np.array(list(buf), dtype=np.int64).reshape(16, 16)
import os
if os.environ.get("EKF_LINE_LOOP", "0") == "1":      # the first form's stamps
    names = ["start", "after landmark loop", "after block min/atomic", "after barrier1", "before row load", "after staging sync", "after rows", "after barrier2"]
    order = list(range(8))
else:                                               # k_scan_lines2
    names = {0: "start", 1: "gate done", 2: "min+atomic", 3: "barrier", 8: "jbest read", 9: "record read", 10: "robot gains", 11: "own loads issued",
             4: "staging sync", 12: "corrections", 13: "own gains", 5: "hot update+stores", 6: "robot block"}
    order = [0, 1, 2, 3, 8, 9, 10, 11, 4, 12, 13, 5, 6]
for line in range(9):
    row = t[line]
    if row[0] == 0: continue
    base = row[0]
    print("line", line, " ".join("%s=%+.2f" % (names[i][:16], (row[i] - base) / 1e3) for i in order if row[i] > 0),
          "| next start %+.2fus" % ((t[line + 1, 0] - base) / 1e3) if t[line + 1, 0] > 0 else "")
