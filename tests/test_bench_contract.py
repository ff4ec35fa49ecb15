"""The bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`) prints ONE JSON
line with the keys the driver reads, times a CPU implementation only, and the product arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("workload,unit", [("extract", "scans/s"), ("room", "steps/s"), ("10k", "steps/s")])
def test_reference_arm_prints_one_contract_line(workload, unit):
    out = _run("--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0 and d["higher_is_better"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["gpu_launches"] == 0
    # both arms print the SAME `config` object (the driver compares them): it is built by one function from the workload
    # and the GPU count alone; what a run measured goes to the own arm's `run` object
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.static_config(workload, 1)


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = _run("--steps", "1", "--warmup", "1")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
