#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, '.')
# same scans through the default build and the sincos build: are the states bit-identical?
import subprocess, json
code = '''
import sys, numpy as np
sys.path.insert(0, ".")
from slam_ros_b200 import EkfFilter, scenario as sc
scn = sc.map_scenario(1000, 200, m=8, seed=3)
f = EkfFilter(capacity_lines=1512)
f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
J = []
for s in range(200):
    rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s]); J.append(j)
y, P, L = f.download_live()
np.savez(sys.argv[1], J=np.array(J), y=y, P=P, pose=pose)
'''
for name, lib in (("a", "slam_ros_b200/libekfcuda.so"), ("b", "slam_ros_b200/libekfcuda_sincos.so")):
    env = dict(os.environ, EKF_LIB=lib)
    subprocess.check_call([sys.executable, "-c", code, "/tmp/sc_%s.npz" % name], env=env)
a = np.load("/tmp/sc_a.npz"); b = np.load("/tmp/sc_b.npz")
print("assoc equal", np.array_equal(a["J"], b["J"]), "y bit-equal", np.array_equal(a["y"], b["y"]), "P bit-equal", np.array_equal(a["P"], b["P"]), "max |dP|", np.abs(a["P"] - b["P"]).max())
PY
for v in "" _sincos; do
EKF_LIB=slam_ros_b200/libekfcuda$v.so timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_1k_sc$v.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_1k_sc$v.json').read().strip().split('\n')[-1]); print('$v 1k value',d['value'],'e2e',d['e2e']['value'])"
EKF_LIB=slam_ros_b200/libekfcuda$v.so timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_sc$v.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_sc$v.json').read().strip().split('\n')[-1]); print('$v mc value',d['value'],'e2e',d['e2e']['value'])"
done
