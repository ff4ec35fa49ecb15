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


# A/B builds of the same sources, loaded through EKF_LIB by tests and measurement scripts (never by default):
#   exact   -DEKF_EXACT_RANK2: the reference's four-rounding p - (ks.x k.x + ks.y k.y) instead of two FMAs
#   timing  -DEKF_LINE_TIMING: %globaltimer stamps of the line loop's phases (scripts/line_timing*.py)
#   mctiming -DEKFB_TIMING: clock64 phase totals of one Monte-Carlo filter's scan, printed by the kernel
VARIANTS = {"exact": ["-DEKF_EXACT_RANK2"], "timing": ["-DEKF_LINE_TIMING"], "mctiming": ["-DEKFB_TIMING"], "mcfull": ["-DEKFB_FULL_PASS"], "sepsincos": ["-DEKF_SEPARATE_SINCOS"]}


def variant_path(name):
    return os.path.join(HERE, "libekfcuda_%s.so" % name)


def _stale(out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(out, extra, verbose=False):
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ))
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


def build(force=False, verbose=False, variants=("exact",)):
    """Builds libekfcuda.so and the A/B variants the tests load (each only when stale)."""
    if force or stale():
        _compile(OUT, os.environ.get("EKF_NVCC_EXTRA", "").split(), verbose)
    for v in variants:
        build_variant(v, force=force)
    return OUT


def build_variant(name, force=False):
    out = variant_path(name)
    if not force and not _stale(out):
        return out
    return _compile(out, VARIANTS[name])


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
