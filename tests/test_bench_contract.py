"""bench.py's JSON contract: the keys the driver reads, for the reference arm (CPU, runs anywhere) and for our arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run_bench(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one line on stdout, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    j = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--streams", "8")
    assert BASE_KEYS <= set(j), BASE_KEYS - set(j)
    assert j["impl"] == "reference" and j["steps"] == 2 and j["warmup"] == 1 and j["unit"] == "scans/s"
    assert j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and "streams" in cb["sample"]
    assert j["gpu_launches"] == 0 and "workload" in j["config"] and "model" not in j["config"]


@pytest.mark.gpu
def test_our_arm_line():
    j = run_bench("--steps", "6", "--warmup", "3", "--streams", "128", "--cpu-seconds", "1")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(j), (BASE_KEYS | {"roofline", "clocks"}) - set(j)
    assert j["impl"] == "ours" and j["n_gpus"] == 1 and j["dtype"] == "f32" and j["scaling"] == "weak"
    assert j["gpu_launches"] == 2 * j["steps"]                        # k_match + k_raycast per step
    assert j["scans"] == 128 * j["steps"] == j["integrated_scans"]    # gate 0/0: every scan matched AND integrated
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 128 * (360 * 8 + 4) and e["d2h_bytes_per_step"] == 128 * 12
    assert e["pose_checksum"] == e["ranges_api"]["pose_checksum"] == e["pipelined_api"]["pose_checksum"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(r["kernels"]) == {"k_match", "k_raycast"} and all(k["avg_launch_ms"] > 0 for k in r["kernels"].values())
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"])
