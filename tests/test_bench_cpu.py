"""bench.py's CPU arm (--impl reference) runs without a GPU and prints the contract's JSON line;
the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("workload, extra", [("c2", ["--res", "48", "--rays", "40000"]), ("c3", ["--res", "32", "--rays", "20000"]),
                                            ("c5", ["--res", "32", "--width", "128", "--height", "64"]), ("c1", [])])
def test_reference_arm_prints_the_contract_line(workload, extra):
    r = _run("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1", *extra)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--workload", "c2", "--res", "16", "--rays", "1000", "--steps", "1")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
