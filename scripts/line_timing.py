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
names = ["start", "after landmark loop", "after block min/atomic", "after barrier1", "before row load", "after staging sync", "after rows", "after barrier2"]
for line in range(9):
    row = t[line, :8]
    if row[0] == 0: continue
    base = row[0]
    print("line", line, " ".join("%s=%+.2fus" % (names[i][:14], (row[i] - base) / 1e3) for i in range(8) if row[i] > 0),
          "| next start %+.2fus" % ((t[line + 1, 0] - base) / 1e3) if t[line + 1, 0] > 0 else "")
