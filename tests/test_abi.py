"""CPU checks of the drop-in boundary: libekfcuda.so loads, exports every symbol include/ekf.h declares
(and nothing of the header is missing from the binding), and refuses to run without a CUDA device."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ekf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ekf_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    from slam_ros_b200.ekf import SYMBOLS
    assert sorted(SYMBOLS) == header_symbols()


def test_library_exports_every_declared_symbol(libekf):
    from slam_ros_b200 import library_path
    out = subprocess.check_output(["nm", "-D", "--defined-only", library_path()]).decode()
    exported = set(re.findall(r" T (ekf_[a-z_0-9]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    for s in header_symbols():
        assert hasattr(libekf, s)
    assert b"sm_100a" in libekf.ekf_version()


def test_library_is_built_for_sm_100a_only():
    from slam_ros_b200 import library_path
    out = subprocess.run(["cuobjdump", "-lelf", library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_device(libekf):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from slam_ros_b200 import EkfFilter, EkfBatch, EkfError
    with pytest.raises(EkfError) as e:
        EkfFilter(capacity_lines=10)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(EkfError):
        EkfBatch(4, capacity_lines=10)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under slam_ros_b200/ may reference it."""
    pkg = os.path.join(ROOT, "slam_ros_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ekforacle" not in txt and "slamref" not in txt, fn


def test_default_config_matches_reference_constants(libekf):
    from slam_ros_b200.ekf import default_config
    cfg = default_config()
    assert (cfg.capacity_lines, cfg.gate, cfg.encoder_noise, cfg.reset_headroom) == (100, 0.4, 0.024, 10)   # Robot.h:13-17, Robot.cpp:893


def test_header_is_plain_c99():
    """include/ekf.h is the boundary a C, Go (cgo), Java (JNI) or Python (ctypes) host binds: it must compile as C99
    with no extension and no C++."""
    import shutil
    import subprocess
    import tempfile
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "abi.c")
        with open(src, "w") as f:
            f.write('#include "ekf.h"\nint main(void) { ekf_config c; ekf_lx* lx = 0; (void)lx; return ekf_default_config(&c) ? 1 : 0; }\n')
        out = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                              "-I", os.path.join(root, "include"), src], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
