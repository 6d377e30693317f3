"""bench.py's reference arm runs the unmodified reference on the host cores (no GPU needed): the
JSON line carries the contract keys.  The B200 arm itself is exercised on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _ref_ready():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libintp_ref_cell.so")) or os.path.isdir("/root/reference")


@pytest.mark.skipif(not _ref_ready(), reason="reference shim not built and /root/reference absent")
def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpts/s" and line["value"] > 0
    assert line["metric"] == "3D cubic fp64 value+gradient eval" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["n_gpus"] == 1 and line["steps"] == 1


@pytest.mark.gpu
def test_b200_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu",
                        "--queries", str(1 << 24), "--no-solve", "--no-other"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["value"] > 0 and line["gpu_launches"] > 0 and line["dtype"] == "f64"
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["peak"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert line["e2e"]["h2d_bytes_per_step"] == (1 << 24) * 24 and line["e2e"]["d2h_bytes_per_step"] == (1 << 24) * 32
