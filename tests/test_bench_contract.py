"""bench.py's output contract: the reference arm on CPU, the GPU arm on a B200 (one JSON line with the keys the driver reads)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert d["impl"] == "reference" and d["metric"] == "resnet18_bbb_mc_images_per_sec_S100" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "2", "--warmup", "3"], 900)
    assert d["metric"] == "resnet18_bbb_mc_images_per_sec_S100" and d["unit"] == "images/s" and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] >= 3 and d["value"] > 1000 and d["dtype"] == "tf32" and d["data"] == "synthetic"
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and "workload" in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 256 * 3 * 32 * 32 * 4 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and 0 < r["frac"] < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["value"] > 0 and c["cores"] >= 1 and c["kind"] in ("reference", "port") and d["value"] / c["value"] > 10
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    m = d["metrics_check"]
    assert 0.0 <= m["error"] <= 1.0 and m["nll"] > 0 and 0.0 <= m["ece"] <= 1.0
