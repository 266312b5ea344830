"""bench.py's output contract: exactly ONE line on stdout, a JSON object carrying the keys the driver reads.
CPU: the reference arm (`--impl reference`: unmodified abPOA on the host cores) on a small slice.
GPU: the product arm on a small slice (device-resident + end-to-end legs)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"}


def _run(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, f"stdout must hold the result line only, got {len(lines)} lines: {r.stdout[:400]!r}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--blocks", "32", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == "poa_dp_inband_gcells_per_s" and d["higher_is_better"] is True
    assert d["data"] == "synthetic" and "workload" in d["config"] and d["vs_baseline"] is None and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_product_arm_line():
    d = _run("--blocks", "300", "--steps", "1", "--warmup", "0", "--no-cpu")
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["metric"] == "poa_dp_inband_gcells_per_s" and d["unit"] == "Gcells/s" and d["n_gpus"] == 1 and d["dtype"] == "int16"
    assert d["data"] == "synthetic" and "workload" in d["config"] and d["vs_baseline"] is None and d["value"] > 0
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "hbm" and rf["unit"] == "GB/s"
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == d["unit"]
    assert d["gpu_launches"] >= 1 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
