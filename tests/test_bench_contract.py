"""bench.py prints ONE JSON line with the keys the driver reads.  The reference arm runs on the CPU (bounded
sample of the workload); the product arm needs a GPU and is exercised on the small config 2."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run("--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= d.keys() and d["impl"] == "reference"
    assert d["metric"] == d["unit"] == "Gcell-steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Gcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_product_arm_json_line():
    d = _run("--workload", "c2", "--steps", "3", "--warmup", "3")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "gpu_launches", "clocks"} <= d.keys() and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["dtype"] == "f32" and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert d["gpu_launches"] >= 2 * d["steps"]                  # projection + advection per step
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= d["clocks"].keys()
