"""The driver's contract of ``bench.py --impl reference`` (the arm that needs no GPU): one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--contigs", "300"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "genes/sec CRF marginal inference" and line["unit"] == "genes/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] >= 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert "workload" in line["config"] and not any(k in line["config"] for k in ("model", "seq_len", "global_batch"))
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["sample"] and cpu["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


import pytest  # noqa: E402


@pytest.mark.gpu
def test_own_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "3", "--warmup", "3", "--contigs", "600",
                          "--no-configs", "--no-sharded"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["gpu_launches"] == 3
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["bit_identical_to_device_path"] is True
    assert 0 < e2e["value"] < line["value"]
    assert line["parity_max_abs_err_vs_oracle"] <= 1e-5 and line["f64"]["parity_max_abs_err_vs_oracle"] <= 1e-12
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
