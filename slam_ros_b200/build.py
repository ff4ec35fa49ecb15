"""Build libekfcuda.so (in-tree) for sm_100a with nvcc.  `python -m slam_ros_b200.build`.

The shared library is the product: include/ekf.h is its whole interface.  It is built next to this file so
that it travels to the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libekfcuda.so")
SOURCES = ["ekf_kernels.cu", "ekf_api.cu", "ekf_batch.cu", "ekf_lines.cu"]
HEADERS = ["ekf_internal.h", "ekf_device.cuh", os.path.join("..", "..", "include", "ekf.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # the gate / gain / sweep arithmetic is un-contracted by contract
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libekfcuda cannot be built (there is no CPU fallback)")
    return p


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    extra = os.environ.get("EKF_NVCC_EXTRA", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT, "-ldl"]
    env = dict(os.environ)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
